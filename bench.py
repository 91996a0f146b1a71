#!/usr/bin/env python3
"""bench.py -- BWT+MTF+RLE+Huffman(+CRC32C) encode/decode throughput on B200 (BASELINE.json metric).

A *step* is one pass of the hot path over one batch of synthetic input: encode every block of the
workload, then decode every block back. At N GPUs every rank owns its own workload of the same size
(blocks are independent: no data-path collective, weak scaling); `value` is the whole-job round-trip
throughput in 10^9 uncompressed bytes per second, max-over-ranks device time.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload all|text|random|periodic|repeat251] [--size-mib M]
    python bench.py --workload c5 --gpus G      # ONE job sharded over G GPUs of this process, no torchrun
    python bench.py --impl reference ...        # the reference's own CPU implementation on the host cores

The headline (top-level keys) is BASELINE.json configs[1]: 1 GiB English-like text, 1 MiB blocks. With
`--workload all` (the default) the same line also carries `per_workload`: text, random (configs[2]),
periodic16_8mib and repeat251_8mib (configs[3]) -- encode/decode GB/s, chain roofline, dominant kernel, parity sample.
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

# name -> (generator kind, block bytes): the BASELINE.json shapes
SHAPES = {"text": ("text", 1 << 20), "random": ("random", 1 << 20), "periodic16_8mib": ("periodic", 8 << 20),
          "repeat251_8mib": ("repeat251", 8 << 20)}
ALIASES = {"periodic": "periodic16_8mib", "repeat251": "repeat251_8mib"}
DESCR = {"text": "synthetic English-like text (lorem vocabulary, splitmix64)", "random": "uniform random bytes (splitmix64)",
         "periodic": "period-16 text (exactly periodic blocks)", "repeat251": "251-byte random pattern repeated (long repeats, not cyclic)"}


def load_vocab():
    with open(os.path.join(ROOT, "tests", "golden", "reference_vectors.json")) as f:
        return json.load(f)["vocab"]


def make_workload(kind, nbytes, seed):
    """C2 text-like, C3 uniform random, C4a exactly periodic, C4b long repeats
    (tools/libbra_gen.so through bra_workloads.py: measurement tooling, not the product library)."""
    import bra_workloads
    return bra_workloads.make(kind, nbytes, seed, load_vocab() if kind == "text" else None)


def workload_name(kind, nbytes, block):
    return f"{nbytes / 2**30:g} GiB {DESCR[kind]}, {block // 1024} KiB blocks"


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


# ------------------------------------------------------------------------------------------ CPU checkers (oracle/, test infrastructure)
def _checker(block):
    """(encode, decode, kind): the reference's own sources (oracle/_ref/libbra_ref.so), else the oracle port."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    if oracle_lib.have_ref():
        impl = oracle_lib.load_ref()
        return impl.encode_block, (lambda h, p, n: impl.decode_block(h, p)), "reference"
    impl = oracle_lib.Oracle()
    return impl.encode_block, (lambda h, p, n: impl.decode_block(h, p, n)), "port"


def cpu_reference_run(kind, block, nblocks, seed, steps=1, warmup=0):
    """Times the reference's own CPU implementation (oracle/_ref/libbra_ref.so: its sources compiled in
    place) -- or, if that library is absent, the oracle port -- on `nblocks` blocks of the workload,
    one block per worker thread at a time over all host cores, encode then decode like
    reference chunks.c:214-238 / :362-397. Returns dict(value, enc_gbs, dec_gbs, cores, kind, sample)."""
    from concurrent.futures import ThreadPoolExecutor
    cores = os.cpu_count() or 1
    enc, dec, kind_name = _checker(block)
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


def parity_sample(kind, block, host, nblk, hdr, pay, crc, payload_stride, want):
    """Bit-exactness of the TIMED run's output (SURVEY.md 8(d) timing protocol): a deterministic sample of `want` blocks,
    evenly spaced over the workload, is compared -- header (primary index, code lengths, sizes), payload bytes and block
    CRC-32C -- with the CPU checker run on the same input blocks. Checker: the reference's own sources for 1 MiB blocks;
    for 8 MiB blocks, where its rotation sort is infeasible, the oracle port (fast BWT pinned against the reference in
    tests/test_oracle.py). Identical input blocks (the exactly periodic shape) share one checker call."""
    import numpy as np
    from concurrent.futures import ThreadPoolExecutor
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    if block <= (1 << 20) and oracle_lib.have_ref():
        enc, name = oracle_lib.load_ref().encode_block, "reference sources (oracle/_ref/libbra_ref.so)"
    else:
        enc, name = oracle_lib.Oracle().encode_block, "oracle port (oracle/liboracle.so, O(n log n) BWT)"
    want = max(1, min(want, nblk))
    ids = sorted({(i * nblk) // want for i in range(want)})
    hv = host.numpy() if hasattr(host, "numpy") else host
    blocks = {b: hv[b * block:(b + 1) * block].tobytes() for b in ids}
    distinct = {}
    for b in ids:
        distinct.setdefault(blocks[b], []).append(b)
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=os.cpu_count() or 1) as ex:
        exp = dict(zip(distinct.keys(), ex.map(enc, distinct.keys())))
    hdr_h = hdr.view(nblk, 268).cpu().numpy()
    crc_h = crc.cpu().numpy().astype(np.uint32)
    bad = []
    for b in ids:
        eh, ep, ec = exp[blocks[b]]
        c = int(hdr_h[b, 264:268].copy().view(np.uint32)[0])
        got = pay[b * payload_stride: b * payload_stride + c].cpu().numpy().tobytes()
        if hdr_h[b].tobytes() != eh or got != ep or int(crc_h[b]) != ec:
            bad.append(b)
    return {"blocks": len(ids), "distinct_inputs": len(distinct), "ok": not bad, "mismatched_blocks": bad[:8], "checker": name,
            "compared": "268-byte header, payload bytes, block CRC-32C of the timed run's output", "seconds": round(time.perf_counter() - t0, 1)}


# algorithmic bytes one launch of each batch-wide kernel family moves, per element of the batch
ALG_PER_ELEM = {"radix_scatter": 16, "radix_scatter_implicit": 12, "radix_hist": 1, "radix_scatter_u8": 5, "mtf_apply": 2, "mtf_summary": 1, "bwt_ranks": 9, "bwt_heads": 9,
                "bwt_prepare": 16, "bwt_finish": 10, "bwt_gather": 6, "crc32c": 1, "ibwt_walk_len": 4, "ibwt_walk_emit": 5}



def bind_near_gpu(torch, local_rank):
    """Run this rank (and therefore allocate its pinned host buffers: first touch) on the CPUs of the NUMA node its GPU
    hangs off, when the box says which one that is and this process may run there. Returns what was done, for the record."""
    info = {"numa_node": None, "bound": False}
    try:
        pr = torch.cuda.get_device_properties(local_rank)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{str(bdf).lower()}/numa_node").read())
        info["numa_node"] = node
        if node < 0:
            return info
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if allowed:
            os.sched_setaffinity(0, allowed)
            info["bound"] = True
            info["cpus"] = len(allowed)
    except (OSError, ValueError, AttributeError):
        pass
    return info


# ------------------------------------------------------------------------------------------ one shape on this rank's GPU
def run_shape(pkg, torch, dist, kind, block, nbytes, batch, steps, warmup, e2e_steps, rank, world, local_rank, peak, parity_blocks, clocks=False):
    """Device-resident timing (CUDA events on the launching stream, max over ranks), parity sample of the timed output,
    end-to-end timing through the host-buffer API. Returns the record on rank 0, None elsewhere."""
    import numpy as np

    def barrier():
        if dist is not None:
            dist.barrier()

    host = torch.from_numpy(make_workload(kind, nbytes, seed=1 + rank)).pin_memory()
    d_in = host.cuda(non_blocking=True)
    nblk = (nbytes + block - 1) // block
    ctx = pkg.Context(local_rank, block, min(batch, nblk))
    enc_out = ctx.alloc_encode_outputs(nblk)
    dec_out = ctx.alloc_decode_outputs(nblk)
    torch.cuda.synchronize()

    def step():
        hdr, pay, crc = ctx.encode_device(d_in, nbytes, outputs=enc_out)
        mid = torch.cuda.Event(enable_timing=True)
        mid.record()
        ctx.decode_device(hdr, pay, nblk, outputs=dec_out)
        return mid

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()

    pkg.prof_reset()
    pkg.prof_enable(True)
    sampler = ClockSampler(local_rank) if clocks else None
    if sampler:
        sampler.start()
    barrier()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    mids = []
    ev[0].record()
    for i in range(steps):
        mids.append(step())
        ev[i + 1].record()
    torch.cuda.synchronize()
    barrier()
    clk = sampler.stop() if sampler else None
    pkg.prof_enable(False)
    prof = pkg.prof_read()
    total_ms = ev[0].elapsed_time(ev[-1])
    enc_ms = sum(ev[i].elapsed_time(mids[i]) for i in range(steps))
    dec_ms = sum(mids[i].elapsed_time(ev[i + 1]) for i in range(steps))

    # ---- correctness of the timed work (outside the timed region) ---------------------------------------
    out, out_len, crc2, status = dec_out
    assert int(status.abs().sum().item()) == 0, "decode reported corrupt blocks"
    assert int(out_len.sum().item()) == nbytes and torch.equal(out[:nbytes], d_in), "round trip mismatch"
    assert torch.equal(crc2, enc_out[2]), "CRC32C of decoded blocks differs from CRC32C of the input blocks"
    par = parity_sample(kind, block, host, nblk, enc_out[0], enc_out[1], enc_out[2], ctx.payload_stride, parity_blocks)
    assert par["ok"], f"timed output differs from the CPU checker on blocks {par['mismatched_blocks']} ({kind})"
    hdr_h = enc_out[0].view(nblk, 268).cpu().numpy()
    r_sum = int(hdr_h[:, 260:264].copy().view(np.uint32).sum())
    c_sum = int(hdr_h[:, 264:268].copy().view(np.uint32).sum())
    stats = ctx.stats()
    del out, out_len, crc2, status, dec_out
    torch.cuda.empty_cache()

    # ---- end to end through the host-buffer API: pinned host input -> .BRa chunk stream in host memory -> back
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
    ctx.close()
    del stream_buf, plain_buf, enc_out, d_in, host, stream, plain
    torch.cuda.empty_cache()

    # ---- reduce over ranks: max time, summed bytes, every rank's parity ---------------------------------------
    t = torch.tensor([total_ms, enc_ms, dec_ms, e2e_s * 1e3, e2e_enc_s * 1e3, (e2e_s - e2e_enc_s) * 1e3], dtype=torch.float64, device="cuda")
    cnt_sum = torch.tensor([par["blocks"]], dtype=torch.int64, device="cuda")
    cnt_ok = torch.tensor([1 if par["ok"] else 0], dtype=torch.int64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt_sum, op=dist.ReduceOp.SUM)
        dist.all_reduce(cnt_ok, op=dist.ReduceOp.MIN)
    total_ms, enc_ms, dec_ms, e2e_ms, e2e_enc_ms, e2e_dec_ms = [float(x) for x in t.tolist()]
    if rank != 0:
        return None
    par = dict(par, blocks=int(cnt_sum.item()), ok=bool(cnt_ok.item()), blocks_this_rank=par["blocks"])
    job_bytes = nbytes * world
    # dominant kernel family of the timed region, from the live CUDA-event brackets
    fam, (fam_launches, fam_ms) = max(prof.items(), key=lambda kv: kv[1][1])
    kernel_ms = sum(ms for _, ms in prof.values())
    elems = min(batch, nblk) * block                      # elements one launch of a batch-wide kernel covers
    share = min(batch, nblk) / nblk
    alg_table = {k: v * elems for k, v in ALG_PER_ELEM.items()}
    alg_table.update({"huf_dec_sync": c_sum * share, "huf_dec_write": (c_sum + r_sum) * share, "huf_pack": (r_sum + c_sum) * share,
                      "rle_enc_emit": elems + r_sum * share, "rle_enc": elems + r_sum * share, "rle_dec_expand": elems + r_sum * share})
    alg_bytes = alg_table.get(fam)
    traffic = None
    try:
        per_elem = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get(fam)
        if isinstance(per_elem, (int, float)):
            traffic = per_elem * elems  # measured DRAM bytes per element (ncu --set full, see the file) x elements per launch
    except OSError:
        pass
    roof = {"bound": "hbm", "kernel": fam, "launches": fam_launches, "avg_launch_ms": fam_ms / max(fam_launches, 1),
            "share_of_kernel_time": fam_ms / kernel_ms if kernel_ms else None, "peak": peak, "unit": "GB/s",
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
    chain_roof = {"encode": {"alg_bytes": alg_enc, "achieved": alg_enc * steps / (enc_ms / 1e3) / 1e9},
                  "decode": {"alg_bytes": alg_dec, "achieved": alg_dec * steps / (dec_ms / 1e3) / 1e9}, "peak": peak, "unit": "GB/s"}
    for k in ("encode", "decode"):
        chain_roof[k]["frac"] = chain_roof[k]["achieved"] / peak
    roof["chain_frac"] = {"encode": chain_roof["encode"]["frac"], "decode": chain_roof["decode"]["frac"]}
    return {"workload": workload_name(kind, nbytes, block), "steps": steps, "warmup": warmup,
            "value": job_bytes * steps / (total_ms / 1e3) / 1e9, "ms_per_step": total_ms / steps,
            "encode_gbs": job_bytes * steps / (enc_ms / 1e3) / 1e9, "decode_gbs": job_bytes * steps / (dec_ms / 1e3) / 1e9,
            "compressed_ratio": (c_sum + 267 * nblk) / nbytes, "rle_ratio": r_sum / nbytes, "bwt_doubling_rounds": stats["bwt_rounds"],
            "huffman_sync_sweeps": stats["huf_sweeps"], "clocks": clk,
            "e2e": {"value": job_bytes / (e2e_ms / 1e3) / 1e9, "unit": UNIT, "steps": e2e_steps, "h2d_bytes_per_step": nbytes + stream_bytes,
                    "d2h_bytes_per_step": stream_bytes + nbytes, "api": "bra_b200_encode_host + bra_b200_decode_host (pinned host buffers)",
                    "encode_gbs": job_bytes / (e2e_enc_ms / 1e3) / 1e9, "decode_gbs": job_bytes / (e2e_dec_ms / 1e3) / 1e9,
                    "pinned_copy_gbs": {"h2d": pc[0], "d2h": pc[1]}},
            "gpu_launches": sum(n for n, _ in prof.values()),
            "kernel_ms": {k: round(v[1] / steps, 3) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1]) if v[0]},
            "roofline": roof, "chain_roofline": chain_roof, "parity_sample": par}


# ------------------------------------------------------------------------------------------ C5: one job over G GPUs of one process
def run_c5(args, emit):
    """16 files (6 text, 6 random, 2 period-16, 2 repeat-251), 1 MiB blocks, packed as ONE job whose block
    list is sharded over the GPUs of this process by bra_b200_pool (dynamic block-range queue, ordered chunk stream, CRC
    chains folded with bra_crc32c_combine); every file's stream and entry CRC chain is compared with a single-GPU run."""
    import numpy as np
    import torch
    import bra_pkg
    pkg = bra_pkg.load()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the compression path has no CPU fallback")
    G = min(args.gpus, torch.cuda.device_count())
    block = 1 << 20
    file_bytes = args.size_mib << 20
    kinds = [("text", 10 + i) for i in range(6)] + [("random", 20 + i) for i in range(6)] + [("periodic", 0), ("periodic", 0), ("repeat251", 0), ("repeat251", 0)]
    kinds = kinds[: args.c5_files] if args.c5_files else kinds
    pool = pkg.Pool(list(range(G)), block, args.c5_range_blocks)
    single = pkg.Context(0, block, 256) if args.c5_check else None
    enc_s = dec_s = 0.0
    total = 0
    per_file = []
    for fi, (kind, seed) in enumerate(kinds):
        host = torch.from_numpy(make_workload(kind, file_bytes, seed)).pin_memory()
        sbuf = torch.empty(pool.encode_bound(file_bytes), dtype=torch.uint8).pin_memory()
        pbuf = torch.empty(file_bytes, dtype=torch.uint8).pin_memory()
        if fi == 0:
            pool.encode_host(host, out=sbuf)  # warm-up: staging buffers of every context
        t0 = time.perf_counter()
        stream, chain = pool.encode_host(host, out=sbuf)
        t1 = time.perf_counter()
        plain, chain2 = pool.decode_host(stream, file_bytes, out=pbuf)
        t2 = time.perf_counter()
        assert chain == chain2 and torch.equal(pbuf, host), f"file {fi}: round trip mismatch"
        ok = None
        if single is not None:
            s1, c1 = single.encode_host(host)
            ok = bool(c1 == chain and s1.nbytes == stream.numel() and np.array_equal(s1, stream.numpy()))
            assert ok, f"file {fi}: the sharded stream differs from the single-GPU stream"
        enc_s += t1 - t0
        dec_s += t2 - t1
        total += file_bytes
        per_file.append({"file": f"f{fi:02d}", "kind": kind, "stream_bytes": int(stream.numel()), "crc_chain": f"{chain:08X}", "encode_s": round(t1 - t0, 3),
                         "decode_s": round(t2 - t1, 3), "equals_single_gpu": ok})
        del host, sbuf, pbuf, stream, plain
    st = pool.stats()
    pool.close()
    if single is not None:
        single.close()
    line = {"metric": METRIC, "value": total / (enc_s + dec_s) / 1e9, "unit": UNIT, "n_gpus": G, "steps": 1, "warmup": 1, "ms_per_step": (enc_s + dec_s) * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": f"C5: {len(kinds)} files x {file_bytes / 2**30:g} GiB (text, random, period-16, repeat-251), 1 MiB blocks, one job sharded over {G} GPUs of one process",
                       "block_bytes": block, "range_blocks": args.c5_range_blocks, "sharding": "dynamic block-range queue over per-GPU contexts, ordered stream, CRC chains folded by combine"},
            "encode_gbs": total / enc_s / 1e9, "decode_gbs": total / dec_s / 1e9,
            "e2e": {"value": total / (enc_s + dec_s) / 1e9, "unit": UNIT, "h2d_bytes_per_step": None, "d2h_bytes_per_step": None,
                    "api": "bra_b200_pool_encode_host + bra_b200_pool_decode_host (pinned host buffers)"},
            "per_gpu_ranges": st, "files": per_file}
    emit(line)


# ------------------------------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="all", choices=["all", "text", "random", "periodic", "repeat251", "periodic16_8mib", "repeat251_8mib", "c5"])
    ap.add_argument("--size-mib", type=int, default=1024)
    ap.add_argument("--block-kib", type=int, default=0, help="override the shape's block size")
    ap.add_argument("--batch", type=int, default=1024, help="blocks per internal batch (bounds device workspace)")
    ap.add_argument("--e2e-steps", type=int, default=0, help="end-to-end iterations of the headline (default: --steps)")
    ap.add_argument("--other-steps", type=int, default=3, help="timed steps of the non-headline shapes of --workload all")
    ap.add_argument("--parity-blocks", type=int, default=64, help="blocks of the timed output compared with the CPU checker, per shape (whole job)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--c5-files", type=int, default=0)
    ap.add_argument("--c5-range-blocks", type=int, default=64)
    ap.add_argument("--c5-check", type=int, default=1)
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
    nbytes = args.size_mib << 20
    wname = ALIASES.get(args.workload, args.workload)
    if wname == "c5":
        if rank == 0:
            run_c5(args, emit)
        return
    head = "text" if wname == "all" else wname
    shapes = list(SHAPES) if wname == "all" else [head]
    hkind, hblock = SHAPES[head]
    if args.block_kib:
        hblock = args.block_kib * 1024
    config = {"workload": workload_name(hkind, nbytes, hblock), "bytes_per_gpu": nbytes, "block_bytes": hblock,
              "batch_blocks": args.batch, "sharding": "independent blocks per GPU, no collective",
              "l2": f"inputs ({nbytes / 2**20:g} MiB per GPU) {'larger' if nbytes > (126 << 20) else 'NOT larger'} than the 126 MB L2; no flush between steps"}

    if args.impl == "reference":
        # rank 0 alone runs the reference's CPU path on the host cores; the other ranks exit without work
        if rank != 0:
            return
        cores = os.cpu_count() or 1
        nblocks = max(cores, 8)
        r = cpu_reference_run(hkind, hblock, nblocks, seed=1, steps=max(args.steps, 1), warmup=min(args.warmup, 1))
        line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u8", "data": "synthetic", "config": config, "encode_gbs": r["encode_gbs"], "decode_gbs": r["decode_gbs"],
                "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
                "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        if wname == "all":
            # the other BASELINE shapes on the reference's CPU path. Its rotation sort is O(n^2 log n) on repeats (reference
            # src/encoders/bra_bwt.h:27-29): the 8 MiB periodic shapes are infeasible, so -- as BASELINE.md section 4 says --
            # the same generators are timed at 4 KiB blocks and the entry says so.
            per = {"text": {k: r[k] for k in ("value", "encode_gbs", "decode_gbs", "sample")}}
            for name, (kind, block) in SHAPES.items():
                if name == "text":
                    continue
                b = block if block <= (1 << 20) else 4096
                rr = cpu_reference_run(kind, b, nblocks, seed=1)
                per[name] = {k: rr[k] for k in ("value", "encode_gbs", "decode_gbs", "sample")}
                if b != block:
                    per[name]["note"] = f"reference BWT infeasible at {block >> 20} MiB blocks of this shape; timed at 4 KiB blocks of the same generator"
            line["per_workload"] = per
        # hygiene: this arm must never map the product library (generators live in tools/libbra_gen.so)
        with open("/proc/self/maps") as f:
            line["product_library_mapped"] = "libbra_b200.so" in f.read()
        assert not line["product_library_mapped"], "the reference arm mapped libbra_b200.so"
        emit(line)
        return

    import torch
    import bra_pkg
    pkg = bra_pkg.load()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the compression path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    numa = bind_near_gpu(torch, local_rank) if world > 1 else {"numa_node": None, "bound": False}
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json, sustained copy)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    parity_blocks = max(8, -(-args.parity_blocks // world))  # per rank; the job-wide sample stays >= --parity-blocks

    recs = {}
    for name in shapes:
        kind, block = SHAPES[name]
        if name == head and args.block_kib:
            block = hblock
        is_head = name == head
        steps = args.steps if is_head else max(1, min(args.steps, args.other_steps))
        warmup = args.warmup if is_head else min(args.warmup, 3)
        e2e_steps = (args.e2e_steps or args.steps) if is_head else 2
        nblk = (nbytes + block - 1) // block
        recs[name] = run_shape(pkg, torch, dist, kind, block, nbytes, min(args.batch, nblk), steps, warmup, max(1, e2e_steps), rank, world, local_rank,
                               peak, parity_blocks, clocks=is_head)

    line = None
    if rank == 0:
        h = recs[head]
        h["roofline"]["peak_source"] = peak_src
        line = {"metric": METRIC, "value": h["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": h["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
                "data": "synthetic", "config": config, "host_placement": numa}
        for k in ("encode_gbs", "decode_gbs", "compressed_ratio", "rle_ratio", "bwt_doubling_rounds", "huffman_sync_sweeps", "clocks", "e2e", "gpu_launches",
                  "kernel_ms", "roofline", "chain_roofline", "parity_sample"):
            line[k] = h[k]
        if wname == "all":
            line["per_workload"] = {name: {k: r[k] for k in ("workload", "steps", "warmup", "value", "encode_gbs", "decode_gbs", "ms_per_step", "compressed_ratio",
                                                             "bwt_doubling_rounds", "huffman_sync_sweeps", "chain_roofline", "parity_sample")}
                                    | {"e2e": {k: r["e2e"][k] for k in ("value", "encode_gbs", "decode_gbs", "steps")},
                                       "dominant_kernel": {k: r["roofline"][k] for k in ("kernel", "avg_launch_ms", "share_of_kernel_time", "achieved", "frac")},
                                       "kernel_ms_top": dict(list(r["kernel_ms"].items())[:6])}
                                    for name, r in recs.items()}
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            r = cpu_reference_run(hkind, hblock, max(cores, 8), seed=1)
            line["cpu_baseline"] = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"],
                                    "encode_gbs": r["encode_gbs"], "decode_gbs": r["decode_gbs"]}
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        emit(line)


if __name__ == "__main__":
    main()
