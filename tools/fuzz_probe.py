#!/usr/bin/env python3
"""Parity triage on the GPU box: replays the case tests/test_gpu_parity.py::test_batched_fuzz_shapes saved in
gpurun_out/fuzz_fail.json with each optional BWT optimisation switched off in turn (BRA_B200_NO_*), as a whole
batch and block by block, and through the per-stage API. Prints one line per configuration."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def child():
    import numpy as np
    import torch
    import bra_pkg
    from oracle_lib import Oracle, RefApi
    pkg = bra_pkg.load()
    o = Oracle()
    d = json.load(open(os.path.join(ROOT, "tests", "golden", "fuzz_case_31337_608.json")))
    data, block = bytes.fromhex(d["data"]), d["block"]
    nblk = (len(data) + block - 1) // block

    def run(buf, mb):
        ctx = pkg.Context(0, block, mb)
        try:
            n = len(buf)
            nb = (n + block - 1) // block
            hdr, pay, crc = ctx.encode_device(torch.frombuffer(bytearray(buf), dtype=torch.uint8).cuda(), n)
            torch.cuda.synchronize()
            h = hdr.cpu().numpy().reshape(nb, 268)
            bad = []
            for b in range(nb):
                eh, ep, ec = o.encode_block(buf[b * block:(b + 1) * block])
                if h[b].tobytes() != eh:
                    bad.append((b, int.from_bytes(h[b, :4].tobytes(), "little"), int.from_bytes(eh[:4], "little")))
            return bad, ctx.stats()
        finally:
            ctx.close()
    out = {"batch": run(data, d["max_batch"]), "batch_mb1": run(data, 1)}
    for b in range(nblk):
        out[f"block{b}_alone"] = run(data[b * block:(b + 1) * block], 1)
    if os.environ.get("PROBE_STAGE_API"):
        api = RefApi(pkg.LIB_PATH)
        for b in range(nblk):
            blk = data[b * block:(b + 1) * block]
            got = api.bwt_encode2(blk)
            exp = o.bwt_encode(blk)
            out[f"stage_api_block{b}"] = (got[1], exp[1], got[0] == exp[0])
    print("PROBE " + json.dumps(out, default=str))


if __name__ == "__main__":
    # one process: the switches are read with getenv at every call
    configs = (("default", {}), ("no_dense", {"BRA_B200_NO_DENSE": "1"}), ("no_alpha", {"BRA_B200_NO_ALPHA": "1"}),
               ("no_finish", {"BRA_B200_NO_FINISH": "1"}), ("no_dense_no_alpha", {"BRA_B200_NO_DENSE": "1", "BRA_B200_NO_ALPHA": "1"}),
               ("all_off", {"BRA_B200_NO_DENSE": "1", "BRA_B200_NO_ALPHA": "1", "BRA_B200_NO_FINISH": "1", "PROBE_STAGE_API": "1"}))
    for name, env in configs:
        for k in ("BRA_B200_NO_DENSE", "BRA_B200_NO_ALPHA", "BRA_B200_NO_FINISH", "PROBE_STAGE_API"):
            os.environ.pop(k, None)
        os.environ.update(env)
        print(name, end=" ", flush=True)
        try:
            child()
        except Exception as ex:  # noqa: BLE001 -- triage tool: report and go on
            print("FAILED:", repr(ex)[:300])
