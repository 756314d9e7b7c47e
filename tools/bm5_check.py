"""tea_bm_5 (4000x4000, 10 steps to convergence): total calc_w/calc_ur calls and final temperature (golden 95.46235158221428,
thesis call count 42297)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from exploringsycl_b200 import TeaLeaf, read_config  # noqa: E402

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
s, st = read_config(os.path.join(root, "tests", "decks", "tea_4000_cg.in"))
if len(sys.argv) > 1:
    s.end_step = int(sys.argv[1])
app = TeaLeaf(s, st)
summ = app.diffuse()
its = [h["iters_a"] for h in app.history]
print("PDL=%s PW=%s iters %s calls %d temp %.17g" % (os.environ.get("TL_PDL", "1"), os.environ.get("TL_PW", "default"), its,
                                                  sum(2 * (i + 1) for i in its) + 0, summ["temp"]), flush=True)
app.close()
