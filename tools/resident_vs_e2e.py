"""diagnostic: prove time through host buffers vs resident inputs, CUDA-event spans on / off.  python tools/resident_vs_e2e.py [n]"""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sipp_b200
from sipp_b200 import _lib
from sipp_b200.sharded import sharded_prove
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
lib = _lib.load()
dA = torch.empty(n * 64, dtype=torch.uint8, device="cuda"); dB = torch.empty(n * 128, dtype=torch.uint8, device="cuda")
_lib.check(lib.sipp_seeded_inputs_device(2, n, dA.data_ptr(), dB.data_ptr()))
A, B = dA.cpu().numpy().tobytes(), dB.cpu().numpy().tobytes()
Ap, Bp = dA.cpu().pin_memory(), dB.cpu().pin_memory()
cp = lambda t: ctypes.cast(t.data_ptr(), ctypes.c_char_p)
def run(name, fn, profile):
    sipp_b200.set_option(_lib.OPT_PROFILE, profile)
    fn(); sipp_b200.stats(reset=True)
    ts = []
    for _ in range(5):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); ts.append((time.perf_counter() - t0) * 1e3)
    st = sipp_b200.stats(reset=True)
    print("%-34s profile=%d  min %.2f median %.2f ms   transcript_exposed %.2f  miller %.2f fe %.2f fold %.2f other %.2f" % (
        name, profile, min(ts), sorted(ts)[2], st["transcript_ms"] / 5, st["miller_ms"] / 5, st["reduce_fe_ms"] / 5, st["fold_ms"] / 5, st["other_ms"] / 5), flush=True)
for rep in range(2):
    run("sipp_prove_native(host bytes)", lambda: sipp_b200.sipp_prove_native(A, B), 0)
    run("sharded_prove(host pinned)", lambda: sharded_prove(cp(Ap), cp(Bp), n, A, B), 0)
    run("sharded_prove(device ptrs)", lambda: sharded_prove(None, None, n, A, B, device_ptrs=(dA.data_ptr(), dB.data_ptr())), 0)
    run("sharded_prove(device ptrs)", lambda: sharded_prove(None, None, n, A, B, device_ptrs=(dA.data_ptr(), dB.data_ptr())), 1)
    run("sharded_prove(host pinned)", lambda: sharded_prove(cp(Ap), cp(Bp), n, A, B), 1)
