"""probe: where does the end-to-end (host buffers) prove spend its time?  Run on the GPU box."""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sipp_b200
from sipp_b200 import _lib
lib = _lib.load(); _lib.require_gpu_once()
n = 4096
A, B = sipp_b200.seeded_inputs(2, n)
plen = lib.sipp_proof_len(n)
out = ctypes.create_string_buffer(384 * plen)
A_pin = torch.frombuffer(bytearray(A), dtype=torch.uint8).pin_memory()
B_pin = torch.frombuffer(bytearray(B), dtype=torch.uint8).pin_memory()
def run(tag, fn, reps=4):
    for i in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize()
        print("%-28s %.1f ms" % (tag, (time.perf_counter() - t0) * 1e3), flush=True)
run("prove_native bytes", lambda: lib.sipp_prove_native(A, n, B, n, out))
run("prove_native pinned", lambda: lib.sipp_prove_native(ctypes.c_char_p(A_pin.data_ptr()), n, ctypes.c_char_p(B_pin.data_ptr()), n, out))
dA = torch.frombuffer(bytearray(A), dtype=torch.uint8).cuda(); dB = torch.frombuffer(bytearray(B), dtype=torch.uint8).cuda()
def resident():
    ctx = sipp_b200.ProverContext(device_ptrs=(dA.data_ptr(), dB.data_ptr()), n=n); ctx.prove(A, B); ctx.close()
run("resident ctx.prove", resident)
def create_only():
    ctx = sipp_b200.ProverContext(A, B); ctx.close()
run("ctx create+destroy (host)", create_only)
t = _lib.TranscriptState(); lib.sipp_transcript_new(ctypes.byref(t))
run("transcript absorb bytes", lambda: lib.sipp_transcript_append_pairs(ctypes.byref(t), A, B, n), 2)
run("transcript absorb pinned", lambda: lib.sipp_transcript_append_pairs(ctypes.byref(t), ctypes.c_char_p(A_pin.data_ptr()), ctypes.c_char_p(B_pin.data_ptr()), n), 2)
print("poseidon backend", lib.sipp_poseidon_backend())
