"""Race hunt: the same Flickr-shaped segment through the engine N times (CUDA-graph replay, side streams, programmatic
dependent launch, pre-wait operand loads), every result compared bit for bit with the first; interleaved with a second
problem so that caches and the graph cache are disturbed.    python profiles/stress_determinism.py [N]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import distill_ref as R
from multimodal_dataset_distillation_b200 import ops

def main(n):
    prs = [R.make_problem(N=100, B=100, K=8, dt=768, d=2304, seed=s, dropout=True) for s in (0, 1)]
    cu = [{k: (v.cuda() if isinstance(v, torch.Tensor) else v) for k, v in pr.items()} for pr in prs]
    ws = [None, None]
    first = [None, None]
    bad = 0
    for it in range(n):
        j = it % 2
        c = cu[j]
        res = ops.unrolled_match(c["theta0"], c["theta_tgt"], c["Y"], c["U"], c["lr"], c["scale"], c["perms"], c["masks"], ws[j])
        ws[j] = res["workspace"]
        cur = {k: res[k].clone() for k in ("out5", "ce", "dY", "dU")}
        if first[j] is None:
            first[j] = cur
        else:
            for k in cur:
                if not torch.equal(cur[k], first[j][k]):
                    bad += 1
                    print(f"iteration {it} problem {j}: {k} differs, max abs diff {float((cur[k] - first[j][k]).abs().max()):.3e}", flush=True)
        if it % 50 == 0:
            torch.cuda.synchronize()
    torch.cuda.synchronize()
    print(f"{n} iterations, {bad} mismatching tensors; loss {float(first[0]['out5'][2]):.6f} / {float(first[1]['out5'][2]):.6f}")
    sys.exit(1 if bad else 0)

if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 400)
