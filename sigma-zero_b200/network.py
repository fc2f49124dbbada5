"""Drop-in for the reference's network.py: policyNN with the same 252-key state_dict (so
`model.load_state_dict(torch.load("supervised_model_best.pt", map_location="cpu"))` works, play.py:25-28)
whose inference forward -- the one the self-play hot path uses, model.eval() as in train_RL.py:213 -- runs on the
hand-written sm_100a kernels of libszb200 (szb_net_forward), not on torch ops, and has no fallback.

`precision` selects the kernel family: "bf16" = tcgen05/TMEM implicit-GEMM tower, "fp32" = SIMT parity path.

In model.train() mode the same modules are evaluated by torch with autograd (batch-statistics BatchNorm): the reference's own
fine-tuning code path (train_RL.py:77-154), kept as the oracle of the CUDA trainer (trainer.py / csrc/train.cu, which is what
train_RL.train_on_records runs on a GPU) and for CPU-only use."""
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

    def forward(self, x):
        """torch path (training only; network.py:66-83)"""
        out = self.relu(self.bn1(self.conv1(x)))
        out = self.bn2(self.conv2(out))
        return self.relu(out + x)


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

    def forward_torch(self, x, inference=False):
        """the reference's forward (network.py:176-192) on torch ops with autograd: the TRAINING path"""
        x = torch.relu(self.norm_layer(self.conv1(x)))
        x = self.resnet_blocks(x)
        p = torch.relu(self.p_norm1(self.conv_p1(x)))
        p = torch.flatten(self.conv_p2(p), start_dim=1)
        v = torch.relu(self.v_norm(self.conv_v1(x)))
        v = torch.relu(self.fc_v1(torch.flatten(v, start_dim=1)))
        v = torch.tanh(self.fc_v2(v))
        if inference:
            p = torch.softmax(p, dim=1)
        return p, v

    def forward(self, x, inference=False):
        """x: [B,119,8,8] 0/1 planes -> (policy [B,4672], value [B,1]) on x's device.
        model.eval(): the CUDA kernels (BatchNorm folded with its running statistics, as in the reference's self-play,
        train_RL.py:213); no gradient, no fallback.  model.train(): torch ops with autograd (fine-tuning)."""
        if self.training:
            return self.forward_torch(x, inference)
        with torch.no_grad():
            return self._forward_cuda(x, inference)

    def _forward_cuda(self, x, inference):
        eng = runtime.get_engine(min_games=int(x.shape[0]))
        runtime.sync_weights(eng, self)
        packed = runtime.pack_planes(x.detach().cpu().numpy())
        pol, val = eng.net_forward(packed, runtime.evaluator_of(self), logits=not inference)
        return torch.from_numpy(pol).to(x.device), torch.from_numpy(val).unsqueeze(1).to(x.device)
