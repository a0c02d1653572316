import sys, numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import test_overlap as T, helpers as H
fx = np.load(T.GOLD); cfg = H.make_config(int(fx["hash_size"]))
models = [H.cuda_model(cfg, T._state(fx, i), train=False) for i in range(2)]
(l1, a, b), (l2, c, d) = T._run(models, fx, "cuda")
print("loss", abs(l1 / float(fx["loss"]) - 1), abs(l2 / float(fx["loss2"]) - 1))
for g, k in ((a, "g_first1"), (b, "g_first2"), (c, "g2_first1"), (d, "g2_first2")):
    print(k, H.rel_err(g, fx[k]), np.abs(g - fx[k]).round(7).tolist())
