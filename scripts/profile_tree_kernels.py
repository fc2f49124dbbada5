"""Short bulk run of the tree / move-generation kernels for ncu: 65,536 trees, hash evaluator, a few simulations,
then a perft(5) (bulk legal move generation).  No network."""
import sys
sys.path.insert(0, ".")
from sigma_zero_b200 import _lib
from sigma_zero_b200.engine import Engine, EVAL_HASH
from tests import util
G = 65536
eng = Engine(max_games=G, max_searches=8)
eng.reset([-1 if g % 2 == 0 else g % 960 for g in range(G)])
eng.search(6, 2.0, True, EVAL_HASH, want_visits=False, want_children=False)
print(eng.stats())
print(eng.perft(util.wire_pos(util.oracle_game(False, -1).board, _lib), 5))
eng.close()
