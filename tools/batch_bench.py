#!/usr/bin/env python
"""Batched instances (BASELINE config 5: `count` independent n = 128 SIPP instances, lock-step on the device).
    python tools/batch_bench.py [count] [n] [steps] [kpg_max]
Prints one JSON line: instances/s resident (inputs and proofs in HBM) and through host buffers, with the CUDA-event time
per kernel class, and checks a few instances against the single-instance prover."""
import ctypes
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import sipp_b200  # noqa: E402
from sipp_b200 import _lib  # noqa: E402


def main():
    count = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    kpg = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    streams = int(sys.argv[5]) if len(sys.argv) > 5 else 0
    lib = _lib.load()
    _lib.require_gpu_once()
    if kpg:
        sipp_b200.set_option(_lib.OPT_BATCH_KPG_MAX, kpg)
    sipp_b200.set_option(_lib.OPT_BATCH_STREAMS, streams)
    total = n * count
    dA = torch.empty(total * 64, dtype=torch.uint8, device="cuda")
    dB = torch.empty(total * 128, dtype=torch.uint8, device="cuda")
    _lib.check(lib.sipp_seeded_inputs_device(5, total, dA.data_ptr(), dB.data_ptr()))
    plen = lib.sipp_proof_len(n)
    dP = torch.empty(count * plen * 384, dtype=torch.uint8, device="cuda")
    A = dA.cpu().numpy().tobytes()
    B = dB.cpu().numpy().tobytes()

    def resident():
        _lib.check(lib.sipp_prove_native_batch_device(dA.data_ptr(), dB.data_ptr(), n, count, dP.data_ptr()))

    out = ctypes.create_string_buffer(count * plen * 384)

    def e2e():
        _lib.check(lib.sipp_prove_native_batch(A, B, n, count, out))

    resident()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        resident()
    torch.cuda.synchronize()
    t_res = (time.perf_counter() - t0) / steps
    # kernel-class times from a profiled pass (event spans on: the library runs one sub-batch so that the spans do not overlap)
    sipp_b200.set_option(_lib.OPT_PROFILE, 1)
    sipp_b200.stats(reset=True)
    for _ in range(steps):
        resident()
    st = sipp_b200.stats(reset=True)
    sipp_b200.set_option(_lib.OPT_PROFILE, 0)
    e2e()
    t0 = time.perf_counter()
    for _ in range(steps):
        e2e()
    t_e2e = (time.perf_counter() - t0) / steps
    proofs = dP.cpu().numpy().tobytes()
    assert proofs == out.raw, "resident and host-buffer proofs differ"
    for j in (0, count // 2, count - 1):
        a, b = A[64 * n * j:64 * n * (j + 1)], B[128 * n * j:128 * n * (j + 1)]
        assert proofs[j * plen * 384:(j + 1) * plen * 384] == b"".join(sipp_b200.sipp_prove_native(a, b)), j
    # batched verifier on the same proofs (host buffers in, one result per instance out)
    res = (ctypes.c_int * count)()
    lib.sipp_verify_native_batch(A, B, n, count, out.raw, plen, res, None, None, None)
    t0 = time.perf_counter()
    _lib.check(lib.sipp_verify_native_batch(A, B, n, count, out.raw, plen, res, None, None, None))
    t_ver = time.perf_counter() - t0
    assert all(r == 0 for r in res), "a valid proof failed the batched verifier"
    loops = count * (3 * n - 2)
    print(json.dumps({"count": count, "n": n, "steps": steps, "kpg_max": kpg or 32, "streams": streams,
                      "resident_s": t_res, "instances_per_s": count / t_res, "pairs_per_s": total / t_res,
                      "e2e_s": t_e2e, "e2e_instances_per_s": count / t_e2e, "verify_batch_s": t_ver, "verified_per_s": count / t_ver,
                      "miller_ms": st["miller_ms"] / steps, "fe_ms": st["reduce_fe_ms"] / steps, "fold_ms": st["fold_ms"] / steps,
                      "other_ms(decode+transcript)": st["other_ms"] / steps, "launches": st["launches"] // steps,
                      "miller_loops_per_s": loops / (st["miller_ms"] / steps * 1e-3),
                      "miller_frac_of_imad_peak(18.5T)": loops / (st["miller_ms"] / steps * 1e-3) * 9008 * 264 / 18.5e12}))


if __name__ == "__main__":
    main()
