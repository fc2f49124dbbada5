"""sigma-zero_b200 -- B200-native self-play hot path of SigmaZero (batched AlphaZero MCTS on chess).

Host side: thin Python over the C ABI in include/szb200.h (libszb200.so, hand-written CUDA for sm_100a).
The reference's module names are mirrored one to one:

    sigma_zero_b200.mcts.MCTS0            <- mcts.py:24-122
    sigma_zero_b200.mctsnode.Node         <- mctsnode.py:7-63 (read-only view of the device tree)
    sigma_zero_b200.chess_tensor          <- chess_tensor.py (ChessTensor, actionsToTensor, actionToTensor, tensorToAction)
    sigma_zero_b200.network.policyNN      <- network.py:89-192
    sigma_zero_b200.sim                   <- sim.py (play_game, generate_training_data)
    sigma_zero_b200.train_RL              <- train_RL.py:156-244 (self-play fan-out + args only)
    sigma_zero_b200.arena                 <- test_update.py:26-83 (new-vs-current promotion match; a "next" row)

`install_dropin()` additionally registers those modules under the reference's flat names (`mcts`,
`mctsnode`, `chess_tensor`, `network`, `sim`) so reference-style scripts run unchanged.
There is no CPU fallback: importing works anywhere, but every compute call needs the CUDA library and a GPU.
"""
__version__ = "0.1.0"


def install_dropin():
    import importlib
    import sys

    for name in ("chess_compat", "chess_tensor", "network", "mctsnode", "mcts", "sim"):
        mod = importlib.import_module("sigma_zero_b200." + name)
        flat = "chess" if name == "chess_compat" else name
        if flat == "chess" and "chess" in sys.modules:
            continue
        sys.modules[flat] = mod
