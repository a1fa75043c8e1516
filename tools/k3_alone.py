"""Development aid: the gradient kernel alone, back to back (c2), against its time inside the K1->K2->K3 stream."""
import sys, torch
sys.path.insert(0, ".")
import monotonic_rnnt_b200 as mr
from monotonic_rnnt_b200 import _lib
lib = _lib.load()
wl = mr.synth.workload("c2"); dev = torch.device("cuda", 0)
acts = torch.empty((wl.rows, wl.V), dtype=torch.float32, device=dev)
_lib.check(lib.mrnnt_synth_uniform(acts.data_ptr(), wl.elements, 0, 0, torch.cuda.current_stream().cuda_stream), "s")
labels = torch.from_numpy(wl.labels).to(dev); T = torch.from_numpy(wl.T).to(dev); S = torch.from_numpy(wl.S).to(dev)
grads = torch.empty_like(acts)
h = mr.LossHandle(acts, labels, T, S, lengths_host=(wl.T, wl.S))
h.enqueue_forward(wl.blank, True)
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
for name, fn in (("K3 alone, back to back", lambda: h.enqueue_backward(grads)),
                 ("K1+K2 alone, back to back", lambda: h.enqueue_forward(wl.blank, True)),
                 ("K1+K2+K3 (enqueue, no host sync)", lambda: h.enqueue(wl.blank, grads))):
    for _ in range(5): fn()
    torch.cuda.synchronize(); e0.record()
    for _ in range(50): fn()
    e1.record(); torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) / 50 * 1e3:.1f} us per iteration")
