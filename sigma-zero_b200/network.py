"""Drop-in for the reference's network.py: policyNN with the same 252-key state_dict (so
`model.load_state_dict(torch.load("supervised_model_best.pt", map_location="cpu"))` works, play.py:25-28)
whose forward runs on the hand-written sm_100a kernels of libszb200 (szb_net_forward), not on torch ops.

The torch modules below are parameter containers only (state_dict / load_state_dict / .to()).  `precision`
selects the kernel family: "bf16" = tcgen05/TMEM implicit-GEMM tower, "fp32" = SIMT parity path."""
import torch
import torch.nn as nn

from . import runtime


class BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, inplanes=256, planes=256):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 3, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(planes, planes, 3, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)

    def forward(self, x):  # pragma: no cover - parameters only
        raise RuntimeError("BasicBlock is a parameter container; policyNN.forward runs the CUDA kernels")


class policyNN(nn.Module):
    def __init__(self, config):
        super().__init__()
        inc = config.get("in_channels", 119)
        if inc != 119:
            raise ValueError("the CUDA network is built for the reference's 119 input planes")
        self.precision = config.get("precision", runtime.DEFAULT_PRECISION)
        self.conv1 = nn.Conv2d(inc, 256, kernel_size=3, padding=1, bias=False)
        self.norm_layer = nn.BatchNorm2d(256)
        self.conv_p1 = nn.Conv2d(256, 256, kernel_size=1, bias=False)
        self.p_norm1 = nn.BatchNorm2d(256)
        self.conv_p2 = nn.Conv2d(256, 73, kernel_size=1)
        self.conv_v1 = nn.Conv2d(256, 1, kernel_size=1, bias=False)
        self.v_norm = nn.BatchNorm2d(1)
        self.fc_v1 = nn.Linear(64, 256)
        self.fc_v2 = nn.Linear(256, 1)
        self.resnet_blocks = nn.Sequential(*[BasicBlock(256, 256) for _ in range(19)])

    @torch.no_grad()
    def forward(self, x, inference=False):
        """x: [B,119,8,8] 0/1 planes (any dtype/device) -> (policy [B,4672], value [B,1]) on x's device.
        BatchNorm runs in eval mode (running statistics), as in the reference's self-play (train_RL.py:213)."""
        if self.training:
            raise RuntimeError("the CUDA forward is inference-only (model.eval()); training is outside the self-play hot path")
        eng = runtime.get_engine(min_games=int(x.shape[0]))
        runtime.sync_weights(eng, self)
        packed = runtime.pack_planes(x.detach().cpu().numpy())
        pol, val = eng.net_forward(packed, runtime.evaluator_of(self), logits=not inference)
        return torch.from_numpy(pol).to(x.device), torch.from_numpy(val).unsqueeze(1).to(x.device)
