#!/usr/bin/env python
"""bench.py -- scanned GB/s of the relative-search hot path on N B200s (one JSON line on rank 0).

A "step" is one pass of the hot path over one batch: every search of the workload (cfg2: the 16-bit
little-endian search and the 16-bit big-endian search) over this rank's slice of the synthetic blob,
followed -- for N > 1 -- by the NCCL gather of the match offsets to rank 0.

    value  : whole-job scanned bytes / second with the blob already resident in HBM (max over ranks)
    e2e    : same metric through the public call with HOST (pinned) buffers: H2D copy + scan + D2H of results
    roofline: the streaming filter kernel (k_filter) against the measured HBM copy bandwidth
    cpu_baseline / --impl reference: the reference's own multithreaded SearchEngine::run (oracle/_ref,
             compiled from the unmodified reference sources) on the host cores of the same box

Weak scaling: every rank scans `size` bytes; the file is n_gpus * size, sharded by whole engine blocks
(contiguous byte ranges + (L-1)*W bytes of overlap), no data-path collective before the scan.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--size-mib", type=int, default=0, help="override the per-GPU blob size (development)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-verify", action="store_true")
    return ap.parse_args()


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 6:
                    self.rows.append(parts)
            except Exception:
                pass
            self.stop_flag.wait(0.1)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": float(self.rows[0][1]),
                "reasons": reasons, "samples": len(self.rows)}


# ------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the unmodified reference engine on the host cores
# ------------------------------------------------------------------------------------------------

def reference_step_seconds(w, blob_path, threads):
    """One step of the workload with mmoore::SearchEngine<T>::run (all searches); -> seconds, matches."""
    from _oracle import Ref
    total, matches = 0.0, 0
    for s in w.searches:
        pat = s.pattern
        t0 = time.perf_counter()
        r = Ref.engine(w.bits, blob_path, keyword=pat.get("keyword"), wildcard=pat.get("wildcard", 0),
                       char_seq=pat.get("char_seq", ()), values=pat.get("values"), big_endian=s.big_endian,
                       threads=threads, block=w.block_size)
        total += time.perf_counter() - t0
        matches += len(r["offsets"])
    return total, matches


def cpu_sample(w, cap_bytes):
    """Bounded sample of the workload for the CPU legs: the first min(size, cap) bytes of the blob."""
    import monkey_moore_b200.workloads as wl
    n = min(w.size, cap_bytes)
    ws = w.scaled(n)
    blob = wl.host_blob(ws)
    d = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
    path = os.path.join(d, "mmoore_bench_%d.bin" % os.getpid())
    blob.tofile(path)
    return ws, path, ("first %d MiB of the %s blob, file in %s, block %d, all searches of a step"
                      % (n >> 20, w.key, d, w.block_size))


def run_reference(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from _oracle import Ref
    kind = "reference" if Ref.available() else "port"
    if kind != "reference":
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libmmref.so missing"}))
        return
    cores = os.cpu_count() or 1
    ws, path, sample = cpu_sample(w, 512 << 20)
    try:
        for _ in range(args.warmup):
            reference_step_seconds(ws, path, cores)
        times = []
        for _ in range(args.steps):
            t, _m = reference_step_seconds(ws, path, cores)
            times.append(t)
    finally:
        os.unlink(path)
    sec = sum(times) / len(times)
    bytes_per_step = ws.size * len(ws.searches)
    gbs = bytes_per_step / sec / 1e9
    line = {"impl": "reference", "metric": "scanned GB/s per search", "value": gbs, "unit": "GB/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u16" if w.bits == 16 else "u8",
            "data": "synthetic",
            "config": {"workload": w.key, "description": w.description, "block_size": w.block_size,
                       "searches": [s.name for s in w.searches]},
            "cpu_baseline": {"value": gbs, "unit": "GB/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": gbs, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------

def run_ours(args, w):
    import torch

    import monkey_moore_b200 as mm
    import monkey_moore_b200.workloads as wl
    from monkey_moore_b200.distributed import shard_bytes

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if mm.device_count() == 0:
        raise SystemExit("bench.py: no CUDA device and no CPU fallback exists for the product path")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    # ---- the blob: this rank's contiguous range of whole blocks (+ overlap) of the n_gpus*size file
    W = w.bits // 8
    total_size = w.size * world
    progs = [mm.Program(w.bits, **s.pattern) for s in w.searches]
    overlap = (max(p.keyword_len for p in progs) - 1) * W
    b0, nb, lo, hi = shard_bytes(total_size, w.block_size, overlap, rank, world)
    b1 = b0 + nb
    blob = wl.device_blob(w, first_byte=lo, nbytes=hi - lo, total_size=total_size)
    assert blob.data_ptr() % 16 == 0
    stream = torch.cuda.current_stream()
    mm.set_stream(stream.cuda_stream, True)
    torch.cuda.synchronize()

    comm = None
    if world > 1:
        def bcast(raw):
            t = torch.tensor(list(raw), dtype=torch.uint8, device="cuda")
            dist.broadcast(t, src=0)
            return bytes(t.cpu().tolist())
        comm = mm.Comm(rank, world, bcast)     # the library's own NCCL communicator for the result gather

    pending_gather = [None]

    def enqueue_step():
        # the searches of a step are independent: all of them are enqueued at once (no host wait in between)
        return [prog.engine_scan(blob, w.block_size, big_endian=s.big_endian, file_size=total_size,
                                 first_block=b0, num_blocks=b1 - b0, asynchronous=True)
                for prog, s in zip(progs, w.searches)]

    def complete_step(held, collect=None):
        launches, filt_ms, filt_bytes = 0, 0.0, 0
        for res in held:
            st = res.stats()
            launches += st["launches"]
            filt_ms += st["ms_filter"]
            filt_bytes += st["bytes_scanned"] + 8 * res.count
        if world > 1:
            # ONE grouped NCCL op per step, only enqueued here; rank 0 completes the previous step's gather
            # (header read, possible spill receives) when its object is dropped below
            gathered = comm.gather(held, lazy=True)
            if collect is not None and gathered is not None:
                collect.extend(torch.from_numpy(g[0].astype(np.int64)) for g in gathered.fetch())
            pending_gather[0] = gathered        # dropping the previous one closes it (and its result lists)
        else:
            if collect is not None:
                collect.extend(r.torch_offsets() for r in held)
            for r in held:
                r.close()
        return launches, filt_ms, filt_bytes

    def step(collect=None):
        return complete_step(enqueue_step(), collect)

    def drain():
        # Rank 0 completes the gather it only enqueued (header read, receives of lists that overflowed the packed
        # buffer).  Must happen before any barrier: the other ranks' overflow sends sit on their streams until rank 0
        # posts the receives, and a barrier behind those sends would wait for a rank 0 that waits in the barrier.
        g = pending_gather[0]
        if g is not None:
            g.counts()

    def barrier():
        drain()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.15)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0
    filt_ms, filt_bytes = 0.0, 0
    barrier()
    e0.record(stream)
    t0 = time.perf_counter()
    # software pipeline over the steps: step k+1 is enqueued before the host collects step k (counts, statistics,
    # gather), so the device never waits for the host between steps; every step is complete before e1 / the barrier
    prev = None
    for _ in range(args.steps):
        cur = enqueue_step()
        if prev is not None:
            l, fm, fb = complete_step(prev)
            launches += l
            filt_ms += fm
            filt_bytes += fb
        prev = cur
    l, fm, fb = complete_step(prev)
    launches += l
    filt_ms += fm
    filt_bytes += fb
    drain()                 # the last step's gather completes inside the timed region
    e1.record(stream)
    barrier()
    wall = time.perf_counter() - t0
    dev_ms = e0.elapsed_time(e1)
    sampler.stop_flag.set()
    sampler.join()
    sec = max(wall, dev_ms / 1e3)
    if world > 1:
        tmax = torch.tensor([sec], dtype=torch.float64, device="cuda")
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        sec = float(tmax.item())
        lt = torch.tensor([launches], dtype=torch.int64, device="cuda")
        dist.all_reduce(lt)
        launches = int(lt.item())
    bytes_per_step = total_size * len(w.searches)
    value = bytes_per_step * args.steps / sec / 1e9

    # ---- roofline leg: the filter kernel ALONE.  Inside the timed region consecutive scans overlap on two streams
    # (the resolve of one beside the filter of the next), so the per-kernel CUDA events there measure kernels that
    # share the memory system.  Here every scan is completed before the next is enqueued; the events sit on the
    # stream the kernel is launched on, around the filter launch only.
    alone_ms, alone_bytes, alone_n = 0.0, 0, 0
    for _ in range(10):
        for prog, s_ in zip(progs, w.searches):
            res = prog.engine_scan(blob, w.block_size, big_endian=s_.big_endian, file_size=total_size,
                                   first_block=b0, num_blocks=b1 - b0)
            st = res.stats()
            alone_ms += st["ms_filter"]
            alone_bytes += st["bytes_scanned"] + 8 * res.count
            alone_n += 1
            res.close()
    torch.cuda.synchronize()

    # ---- e2e: host (pinned) buffers through the public call, H2D + scan + D2H of the results
    mm.set_stream(None, False)
    host = torch.empty(hi - lo, dtype=torch.uint8).pin_memory()
    host.copy_(blob)
    torch.cuda.synchronize()
    host_np = host.numpy()
    e2e_steps = max(2, min(args.steps, 5))
    d2h = 0

    def e2e_step():
        nonlocal d2h
        d2h = 0
        for prog, s in zip(progs, w.searches):
            res = prog.engine_scan(host_np, w.block_size, big_endian=s.big_endian, file_size=total_size,
                                   first_block=b0, num_blocks=b1 - b0)
            off, val = res.arrays()
            d2h += off.nbytes + val.nbytes // 2
            res.close()
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_sec = time.perf_counter() - t0
    if world > 1:
        tmax = torch.tensor([e2e_sec], dtype=torch.float64, device="cuda")
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        e2e_sec = float(tmax.item())
    e2e_value = bytes_per_step * e2e_steps / e2e_sec / 1e9

    # ---- parity at full size (rank 0, N = 1): the oracle on the same bytes, after the timed region
    parity = None
    if rank == 0 and world == 1 and not args.no_verify:
        from _oracle import Oracle
        got = []
        mm.set_stream(stream.cuda_stream, True)
        step(collect=got)
        mm.set_stream(None, False)
        ok = True
        for prog, s, g in zip(progs, w.searches, got):
            pat = s.pattern
            o = Oracle(w.bits, keyword=pat.get("keyword"), wildcard=pat.get("wildcard", 0),
                       char_seq=pat.get("char_seq", ()), values=pat.get("values"))
            exp, _ = o.engine(host_np, w.block_size, big_endian=s.big_endian, wrap32=False)
            ok = ok and g.cpu().numpy().astype(np.uint64).tolist() == exp.tolist()
        parity = {"checked": "oracle on the full blob, all searches", "bit_exact": bool(ok),
                  "matches": int(sum(int(g.numel()) for g in got))}

    # ---- CPU baseline beside it (rank 0, N = 1 only): the reference engine on a bounded sample
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            from _oracle import Ref
            if Ref.available():
                cores = os.cpu_count() or 1
                ws, path, sample = cpu_sample(w, 512 << 20)
                try:
                    reference_step_seconds(ws, path, cores)
                    t, _m = reference_step_seconds(ws, path, cores)
                finally:
                    os.unlink(path)
                cpu = {"value": ws.size * len(ws.searches) / t / 1e9, "unit": "GB/s", "cores": cores,
                       "kind": "reference", "sample": sample}
        except Exception as e:  # the baseline is reported, never required
            cpu = {"value": None, "unit": "GB/s", "cores": os.cpu_count(), "kind": "reference", "sample": "failed: %s" % e}

    if rank == 0:
        peak, peak_src = measured_peak()
        n_filter = args.steps * len(w.searches)
        achieved = (alone_bytes / alone_n) / (alone_ms / alone_n) / 1e6       # GB/s, per-launch averages, kernel alone
        achieved_overlapped = (filt_bytes / n_filter) / (filt_ms / n_filter) / 1e6   # same events inside the timed region
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                traffic = json.load(f).get(w.key)
        except Exception:
            pass
        line = {"metric": "scanned GB/s per search", "value": value, "unit": "GB/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": sec / args.steps * 1e3,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u16" if w.bits == 16 else "u8", "data": "synthetic",
                "config": {"workload": w.key, "description": w.description, "bytes_per_gpu": w.size,
                           "block_size": w.block_size, "searches": [s.name for s in w.searches],
                           "l2": "input (%d MiB per GPU) larger than the 126 MB L2, no flush needed" % (w.size >> 20),
                           "parallelism": "block-sharded x%d" % world},
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                             "frac": achieved / peak, "frac_of_nominal_8TBs": achieved / 8000.0, "traffic": traffic, "kernel": "k_filter" if w.bits == 16 else "k_filter8", "peak_source": peak_src,
                             "timing": "CUDA events around the filter launch on its stream, %d launches run alone "
                                       "after the timed region (inside it scans overlap on two streams)" % alone_n,
                             "achieved_in_timed_region": achieved_overlapped,
                             "pipeline_frac": value / world / peak},
                "cpu_baseline": cpu,
                "e2e": {"value": e2e_value, "unit": "GB/s", "h2d_bytes_per_step": int((hi - lo) * len(w.searches)),
                        "d2h_bytes_per_step": int(d2h)},
                "gpu_launches": launches, "clocks": sampler.summary(), "parity": parity,
                "device_ms_per_step": dev_ms / args.steps}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    import monkey_moore_b200.workloads as wl
    w = wl.WORKLOADS[args.workload]
    if args.size_mib:
        w = w.scaled(args.size_mib << 20)
    if args.impl == "reference":
        run_reference(args, w)
    else:
        run_ours(args, w)


if __name__ == "__main__":
    main()
