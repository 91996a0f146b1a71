#!/usr/bin/env python3
"""bench.py -- BWT+MTF+RLE+Huffman(+CRC32C) encode/decode throughput on B200 (BASELINE.json metric).

A *step* is one pass of the hot path over one batch of synthetic input: encode every block of the
workload, then decode every block back. At N GPUs every rank owns its own workload of the same size
(blocks are independent: no data-path collective, weak scaling); `value` is the whole-job round-trip
throughput in 10^9 uncompressed bytes per second, max-over-ranks device time.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload text|random|periodic] [--size-mib M] [--block-kib B]
    python bench.py --impl reference ...   # the reference's own CPU implementation on the host cores

One JSON line on stdout (rank 0). See DESIGN.md "Measurement" for every field.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "BWT+MTF+RLE+Huffman encode/decode GB/s"
UNIT = "GB/s (1e9 uncompressed bytes/s, encode+decode round trip)"


def load_vocab():
    with open(os.path.join(ROOT, "tests", "golden", "reference_vectors.json")) as f:
        return json.load(f)["vocab"]


def make_workload(kind, nbytes, seed):
    """C2 text-like, C3 uniform random, C4a exactly periodic, C4b long repeats
    (tools/libbra_gen.so through bra_workloads.py: measurement tooling, not the product library)."""
    import bra_workloads
    return bra_workloads.make(kind, nbytes, seed, load_vocab() if kind == "text" else None)


def workload_name(kind, nbytes, block):
    names = {"text": "synthetic English-like text (lorem vocabulary, splitmix64)", "random": "uniform random bytes (splitmix64)",
             "periodic": "period-16 text", "repeat251": "251-byte random pattern repeated (long repeats, not cyclic)"}
    return f"{nbytes / 2**30:g} GiB {names[kind]}, {block // 1024} KiB blocks"


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    FIELDS = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, device):
        self.device = device
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.device}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], 0, set()
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 8:
                continue
            try:
                sm.append(float(p[1]))
                mx = max(mx, float(p[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_reference_run(kind, block, nblocks, seed, steps=1, warmup=0):
    """Times the reference's own CPU implementation (oracle/_ref/libbra_ref.so: its sources compiled in
    place) -- or, if that library is absent, the oracle port -- on `nblocks` blocks of the workload,
    one block per worker thread at a time over all host cores, encode then decode like
    reference chunks.c:214-238 / :362-397. Returns dict(value, enc_gbs, dec_gbs, cores, kind, sample)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from concurrent.futures import ThreadPoolExecutor
    import oracle_lib
    cores = os.cpu_count() or 1
    if oracle_lib.have_ref():
        impl, kind_name = oracle_lib.load_ref(), "reference"
        enc = impl.encode_block
        dec = lambda h, p, n: impl.decode_block(h, p)
    else:
        impl, kind_name = oracle_lib.Oracle(), "port"
        enc = impl.encode_block
        dec = lambda h, p, n: impl.decode_block(h, p, n)
    data = make_workload(kind, nblocks * block, seed).tobytes()
    blocks = [data[i * block:(i + 1) * block] for i in range(nblocks)]
    t_enc, t_dec = [], []
    with ThreadPoolExecutor(max_workers=cores) as ex:
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            encoded = list(ex.map(enc, blocks))
            t1 = time.perf_counter()
            plain = list(ex.map(lambda e_b: dec(e_b[0][0], e_b[0][1], len(e_b[1])), zip(encoded, blocks)))
            t2 = time.perf_counter()
            assert all(p == b for p, b in zip(plain, blocks)), "CPU reference round trip failed"
            if it >= warmup:
                t_enc.append(t1 - t0)
                t_dec.append(t2 - t1)
    nbytes = nblocks * block
    te, td = sum(t_enc) / len(t_enc), sum(t_dec) / len(t_dec)
    return {"value": nbytes / (te + td) / 1e9, "encode_gbs": nbytes / te / 1e9, "decode_gbs": nbytes / td / 1e9, "unit": UNIT, "cores": cores,
            "kind": kind_name, "sample": f"{nblocks} blocks x {block} B of the same workload, one block per thread over {cores} threads",
            "ms_per_step": (te + td) * 1e3}


# ------------------------------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="text", choices=["text", "random", "periodic", "repeat251"])
    ap.add_argument("--size-mib", type=int, default=1024)
    ap.add_argument("--block-kib", type=int, default=1024)
    ap.add_argument("--batch", type=int, default=1024, help="blocks per internal batch (bounds device workspace)")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    # stdout carries exactly one JSON line: anything libraries print meanwhile (e.g. NCCL's version banner) goes to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        sys.stdout.flush()
        os.dup2(json_fd, 1)
        print(json.dumps(line), flush=True)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    block = args.block_kib * 1024
    nbytes = args.size_mib << 20
    config = {"workload": workload_name(args.workload, nbytes, block), "bytes_per_gpu": nbytes, "block_bytes": block,
              "batch_blocks": args.batch, "sharding": "independent blocks per GPU, no collective", "l2": "inputs (>= 1 GiB) larger than L2"}

    if args.impl == "reference":
        # rank 0 alone runs the reference's CPU path on the host cores; the other ranks exit without work
        if rank != 0:
            return
        cores = os.cpu_count() or 1
        nblocks = max(cores, 8)
        r = cpu_reference_run(args.workload, block, nblocks, seed=1, steps=max(args.steps, 1), warmup=min(args.warmup, 1))
        line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u8", "data": "synthetic", "config": config, "encode_gbs": r["encode_gbs"], "decode_gbs": r["decode_gbs"],
                "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
                "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        emit(line)
        return

    import numpy as np
    import torch
    import bra_pkg
    pkg = bra_pkg.load()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the compression path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()

    # ---- workload: generated on the host (pinned), resident in HBM before the timed region -------------
    host = torch.from_numpy(make_workload(args.workload, nbytes, seed=1 + rank)).pin_memory()
    d_in = host.cuda(non_blocking=True)
    nblk = (nbytes + block - 1) // block
    ctx = pkg.Context(local_rank, block, min(args.batch, nblk))
    enc_out = ctx.alloc_encode_outputs(nblk)
    dec_out = ctx.alloc_decode_outputs(nblk)
    torch.cuda.synchronize()

    def step():
        hdr, pay, crc = ctx.encode_device(d_in, nbytes, outputs=enc_out)
        mid = torch.cuda.Event(enable_timing=True)
        mid.record()
        ctx.decode_device(hdr, pay, nblk, outputs=dec_out)
        return mid

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    # decode hints (upper bounds of orig_size / encoded_size) are not passed: the device path sizes its grids from the caps

    pkg.prof_reset()
    pkg.prof_enable(True)
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    mids = []
    ev[0].record()
    for i in range(args.steps):
        mids.append(step())
        ev[i + 1].record()
    torch.cuda.synchronize()
    barrier()
    clocks = sampler.stop()
    pkg.prof_enable(False)
    prof = pkg.prof_read()
    total_ms = ev[0].elapsed_time(ev[-1])
    enc_ms = sum(ev[i].elapsed_time(mids[i]) for i in range(args.steps))
    dec_ms = sum(mids[i].elapsed_time(ev[i + 1]) for i in range(args.steps))

    # ---- correctness of the timed work (outside the timed region): exact round trip + sizes -------------
    out, out_len, crc2, status = dec_out
    assert int(status.abs().sum().item()) == 0, "decode reported corrupt blocks"
    assert int(out_len.sum().item()) == nbytes and torch.equal(out[:nbytes], d_in), "round trip mismatch"
    assert torch.equal(crc2, enc_out[2]), "CRC32C of decoded blocks differs from CRC32C of the input blocks"
    hdr_h = enc_out[0].view(nblk, 268).cpu().numpy()
    r_sum = int(hdr_h[:, 260:264].copy().view(np.uint32).sum())
    c_sum = int(hdr_h[:, 264:268].copy().view(np.uint32).sum())
    stats = ctx.stats()

    # ---- end to end through the host-buffer API: pinned host input -> .BRa chunk stream in host memory -> back
    e2e_steps = max(1, args.e2e_steps)
    stream_buf = torch.empty(int(ctx.L.bra_b200_encode_bound(ctx.handle, nbytes)), dtype=torch.uint8).pin_memory()
    plain_buf = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    stream, chain = ctx.encode_host(host, out=stream_buf)       # warm-up (allocates the staging buffers)
    ctx.decode_host(stream, nbytes, out=plain_buf)
    barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e2e_enc_s = 0.0
    for _ in range(e2e_steps):
        t1 = time.perf_counter()
        stream, chain = ctx.encode_host(host, out=stream_buf)   # returns when the stream is complete in host memory
        e2e_enc_s += time.perf_counter() - t1
        plain, chain2 = ctx.decode_host(stream, nbytes, out=plain_buf)
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    e2e_enc_s /= e2e_steps
    assert chain == chain2 and torch.equal(plain_buf, host), "end-to-end round trip mismatch"
    stream_bytes = int(stream.numel())
    # pinned-memory copy rates of this box, to read the end-to-end figure against (outside every timed region)
    pc = []
    for src, dst in ((host, d_in), (d_in, plain_buf)):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        e0.record()
        dst.copy_(src, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        pc.append(nbytes / (e0.elapsed_time(e1) / 1e3) / 1e9)

    # ---- reduce over ranks: max time, summed bytes ---------------------------------------------------------
    t = torch.tensor([total_ms, enc_ms, dec_ms, e2e_s * 1e3, e2e_enc_s * 1e3, (e2e_s - e2e_enc_s) * 1e3], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, enc_ms, dec_ms, e2e_ms, e2e_enc_ms, e2e_dec_ms = [float(x) for x in t.tolist()]
    job_bytes = nbytes * world
    value = job_bytes * args.steps / (total_ms / 1e3) / 1e9

    line = None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json, sustained copy)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        # dominant kernel family of the timed region, from the live CUDA-event brackets
        fam, (fam_launches, fam_ms) = max(prof.items(), key=lambda kv: kv[1][1])
        kernel_ms = sum(ms for _, ms in prof.values())
        elems = min(args.batch, nblk) * block                      # elements one launch of a batch-wide kernel covers
        # algorithmic bytes one launch of each batch-wide kernel family moves; payload-sized
        # kernels use the measured payload / RLE sizes of the batch
        share = min(args.batch, nblk) / nblk
        alg_table = {"radix_scatter": 16 * elems, "radix_hist": 4 * elems, "radix_scatter_u8": 5 * elems, "mtf_apply": 2 * elems,
                     "mtf_summary": elems, "bwt_ranks": 9 * elems, "bwt_heads": 9 * elems, "bwt_prepare": 16 * elems, "bwt_finish": 10 * elems,
                     "bwt_gather": 6 * elems, "crc32c": elems, "ibwt_walk_len": 4 * elems, "ibwt_walk_emit": 5 * elems,
                     "huf_dec_sync": c_sum * share, "huf_dec_write": (c_sum + r_sum) * share, "huf_pack": (r_sum + c_sum) * share,
                     "rle_enc_emit": elems + r_sum * share, "rle_dec_expand": elems + r_sum * share}
        alg_bytes = alg_table.get(fam)
        traffic = None
        try:
            per_elem = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get(fam)
            if isinstance(per_elem, (int, float)):
                traffic = per_elem * elems  # measured DRAM bytes per element (ncu --set full, see the file) x elements per launch
        except OSError:
            pass
        roof = {"bound": "hbm", "kernel": fam, "launches": fam_launches, "avg_launch_ms": fam_ms / max(fam_launches, 1),
                "share_of_kernel_time": fam_ms / kernel_ms if kernel_ms else None, "peak": peak, "unit": "GB/s", "peak_source": peak_src,
                "traffic": traffic, "traffic_source": "profiles/ncu_traffic.json (ncu dram__bytes per element x elements per launch)" if traffic else None}
        if alg_bytes:
            roof["alg_bytes_per_launch"] = alg_bytes
            roof["achieved"] = alg_bytes / (roof["avg_launch_ms"] / 1e3) / 1e9
            roof["frac"] = roof["achieved"] / peak
        else:
            roof["achieved"] = None
            roof["frac"] = None
        # whole-chain algorithmic traffic: ALG_ENC = 15n+3r+c+268, ALG_DEC = 16n+2r+c+268 per block
        alg_enc = 15 * nbytes + 3 * r_sum + c_sum + 268 * nblk
        alg_dec = 16 * nbytes + 2 * r_sum + c_sum + 268 * nblk
        chain_roof = {"encode": {"alg_bytes": alg_enc, "achieved": alg_enc * args.steps / (enc_ms / 1e3) / 1e9},
                      "decode": {"alg_bytes": alg_dec, "achieved": alg_dec * args.steps / (dec_ms / 1e3) / 1e9}, "peak": peak, "unit": "GB/s"}
        for k in ("encode", "decode"):
            chain_roof[k]["frac"] = chain_roof[k]["achieved"] / peak
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
                "data": "synthetic", "config": config,
                "encode_gbs": job_bytes * args.steps / (enc_ms / 1e3) / 1e9, "decode_gbs": job_bytes * args.steps / (dec_ms / 1e3) / 1e9,
                "compressed_ratio": (c_sum + 267 * nblk) / nbytes, "rle_ratio": r_sum / nbytes, "bwt_doubling_rounds": stats["bwt_rounds"],
                "huffman_sync_sweeps": stats["huf_sweeps"],
                "clocks": clocks,
                "e2e": {"value": job_bytes / (e2e_ms / 1e3) / 1e9, "unit": UNIT, "h2d_bytes_per_step": nbytes + stream_bytes,
                        "d2h_bytes_per_step": stream_bytes + nbytes, "api": "bra_b200_encode_host + bra_b200_decode_host (pinned host buffers)",
                        "encode_gbs": job_bytes / (e2e_enc_ms / 1e3) / 1e9, "decode_gbs": job_bytes / (e2e_dec_ms / 1e3) / 1e9,
                        "pinned_copy_gbs": {"h2d": pc[0], "d2h": pc[1]}},
                "gpu_launches": sum(n for n, _ in prof.values()),
                "kernel_ms": {k: round(v[1] / args.steps, 3) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1]) if v[0]},
                "roofline": roof, "chain_roofline": chain_roof}
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            r = cpu_reference_run(args.workload, block, max(cores, 8), seed=1)
            line["cpu_baseline"] = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"],
                                    "encode_gbs": r["encode_gbs"], "decode_gbs": r["decode_gbs"]}
    ctx.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        emit(line)


if __name__ == "__main__":
    main()
