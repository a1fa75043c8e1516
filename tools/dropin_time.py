"""Development aid: per-call time of the reference's UNMODIFIED torch binding compiled against our include/
(tests/dropin/_build/monotonic_rnnt_cpp.so): new manager + create_workspace + cost_and_grad + free_workspace per call,
exactly as pytorch_binding/monotonic_rnnt.cu:79-113 does it.  c2 by default."""
import importlib.util, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import monotonic_rnnt_b200 as mr
from monotonic_rnnt_b200 import _lib

so = os.path.join(ROOT, "tests", "dropin", "_build", "monotonic_rnnt_cpp.so")
spec = importlib.util.spec_from_file_location("monotonic_rnnt_cpp", so)
mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
wl = mr.synth.workload(sys.argv[1] if len(sys.argv) > 1 else "c2")
dev = torch.device("cuda", 0)
acts = torch.empty((wl.rows, wl.V), dtype=torch.float32, device=dev)
_lib.check(_lib.load().mrnnt_synth_uniform(acts.data_ptr(), wl.elements, 0, 0, torch.cuda.current_stream().cuda_stream), "s")
labels = torch.from_numpy(wl.labels).to(dev); T = torch.from_numpy(wl.T).to(dev); S = torch.from_numpy(wl.S).to(dev)
grads = torch.empty_like(acts); costs = torch.zeros(wl.B, dtype=torch.float32)
if wl.alignment is not None:   # the align-restricted entry (monotonic_rnnt.cu:116-150)
    al = torch.from_numpy(wl.alignment).to(dev)
    call = lambda: mod.gpu_monotonic_rnnt_align_restrict(acts, labels, T, S, al, wl.max_shift, costs, grads, wl.blank, 0)
else:
    call = lambda: mod.gpu_monotonic_rnnt(acts, labels, T, S, costs, grads, wl.blank, 0)
for _ in range(5):
    assert call() == 0
torch.cuda.synchronize(); t0 = time.perf_counter()
n = 50
for _ in range(n):
    call()
torch.cuda.synchronize()
ms = (time.perf_counter() - t0) / n * 1e3
print(f"{wl.name}: reference torch binding on these headers: {ms:.3f} ms per call ({wl.B / ms * 1e3:.0f} utt/s), cost[0]={costs[0].item():.4f}")
