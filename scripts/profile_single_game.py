"""ncu target: a short search of ONE game (the reference's real API, mcts.py:39 / play.py:40-43): every step is k_tree_step + the
cluster-resident tower k_tower_cl.  Usage: python scripts/profile_single_game.py [games=1] [sims=24]"""
import sys
import torch
sys.path.insert(0, ".")
from oracle import ref_path
from sigma_zero_b200.engine import Engine, EVAL_NET_BF16
G = int(sys.argv[1]) if len(sys.argv) > 1 else 1
S = int(sys.argv[2]) if len(sys.argv) > 2 else 24
torch.manual_seed(0)
eng = Engine(max_games=G, max_searches=S)
eng.load_state_dict(ref_path.build_policy_nn().eval().state_dict())
eng.reset([-1] * G)
v, _, _ = eng.search(S, 2.0, True, EVAL_NET_BF16)
print(v.sum(axis=1), eng.stats())
eng.close()
