import sys, torch, numpy as np
sys.path.insert(0, ".")
import monotonic_rnnt_b200 as mr
from monotonic_rnnt_b200 import _lib
lib = _lib.load()
for name in ("c2", "c3"):
    wl = mr.synth.workload(name)
    dev = torch.device("cuda", 0)
    acts = torch.empty((wl.rows, wl.V), dtype=torch.float32, device=dev)
    _lib.check(lib.mrnnt_synth_uniform(acts.data_ptr(), wl.elements, 0, 0, torch.cuda.current_stream().cuda_stream), "s")
    labels = torch.from_numpy(wl.labels).to(dev); T = torch.from_numpy(wl.T).to(dev); S = torch.from_numpy(wl.S).to(dev)
    grads = torch.empty_like(acts); costs = torch.empty(wl.B, dtype=torch.float32).pin_memory()
    h = mr.LossHandle(acts, labels, T, S, lengths_host=(wl.T, wl.S))
    ref = None
    for pdl in (0, 1, 0, 1):
        h.set_option(_lib.OPT_PDL, pdl)
        for _ in range(5): h.cost_and_grad(wl.blank, grads, costs)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(50): h.cost_and_grad(wl.blank, grads, costs)
        e1.record(); torch.cuda.synchronize()
        g = grads.sum().item(); c = costs.clone()
        if ref is None: ref = (g, c)
        assert g == ref[0] and torch.equal(c, ref[1]), (g, ref[0])
        print(f"{name} pdl={pdl}: {e0.elapsed_time(e1) / 50 * 1e3:.1f} us per call", flush=True)
    h.close()
