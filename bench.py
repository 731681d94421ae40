#!/usr/bin/env python
"""bench.py -- scanned GB/s of the relative-search hot path on N B200s (one JSON line on rank 0).

A "step" is one pass of the hot path over one batch: every search of the workload (cfg2: the 16-bit
little-endian search and the 16-bit big-endian search) over this rank's slice of the synthetic blob,
followed -- for N > 1 -- by the NCCL gather of the match lists to rank 0.

    value    : whole-job scanned bytes / second with the blob already resident in HBM (max over ranks)
    e2e      : same metric through the public call with HOST (pinned) buffers: H2D copy + scan + D2H of results
               (+ the gather for N > 1)
    roofline : the streaming filter kernel (k_filter / k_filter8) against the measured HBM copy bandwidth
    parity   : the result lists against the CPU oracle at FULL size: every rank checks its own shard (match count +
               order-sensitive digest of (offset, table values); the full lists when they are short), rank 0 checks
               the gathered whole against the rank-order composition of the per-rank oracle digests
    cpu_baseline / --impl reference: the reference's own multithreaded SearchEngine::run (oracle/_ref, compiled from
               the unmodified reference sources) on the host cores of the same box, at the struct-default block
               (524288) and at the GUI's block (8 MiB)
    per_config: the other BASELINE configurations with the same fields (N = 1: cfg1, cfg3 at full size, cfg4 / cfg5 at
               their per-GPU slice; N > 1: cfg4 strong-scaled (16 GiB in total) and cfg5 (8 GiB per GPU))
    single_chain: SURVEY 8f-4 -- the cfg4 bytes searched as ONE MonkeyMoore<uint16_t>::search chain: one call at N = 1
               (and four slices through mmg_chain_*), one slice per rank through mmg_comm_search at N > 1 (the slice
               maps cross the ranks in a 128-byte all-gather, the only exchange step of the whole path), checked against
               the oracle's chain slice by slice

Scaling: `weak` -- every rank scans `size` bytes, the file is n_gpus * size; `strong` -- the file is `size` bytes in
total.  Either way the file is sharded by whole engine blocks (contiguous byte ranges + (L-1)*W bytes of overlap) and no
collective touches the data path.
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

METRIC = "scanned GB/s per search"
L2_BYTES = 126 << 20
GUI_BLOCK = 8 << 20          # src/gui/monkey_prefs.cpp:26


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--scaling", default="auto", choices=["auto", "weak", "strong"])
    ap.add_argument("--size-mib", type=int, default=0, help="override the blob size (development)")
    ap.add_argument("--per-config", default="auto", help="auto | none | comma separated workload keys")
    ap.add_argument("--chain", default="auto", choices=["auto", "on", "off"],
                    help="single_chain object (cfg4 bytes as ONE search() chain, SURVEY 8f-4): auto = with the per_config run")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-verify", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
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


def fits_in_host_ram(nbytes):
    try:
        import psutil
        return nbytes < 0.4 * psutil.virtual_memory().available
    except Exception:
        return nbytes <= (8 << 30)


def dtype_of(w):
    return "u16" if w.bits == 16 else "u8"


def resolve_scaling(args, w, world):
    """-> (scaling, bytes this run's file holds in total, bytes per GPU)"""
    scaling = w.scaling if args.scaling == "auto" else args.scaling
    size = (args.size_mib << 20) if args.size_mib else w.size
    if scaling == "strong":
        total = size
        if world == 1 and w.single_gpu_size and not args.size_mib:
            total = w.single_gpu_size          # a multi-GPU configuration measured on one GPU: its per-GPU slice
        return scaling, total, total // world
    return scaling, size * world, size


def config_of(w, scaling, total, world):
    """The `config` object -- identical in both arms (the driver compares them)."""
    return {"workload": w.key, "description": w.description, "block_size": w.block_size,
            "searches": [s.name for s in w.searches], "scaling": scaling, "file_bytes": int(total),
            "bytes_per_gpu": int(total // world), "generator": w.generator}


def pattern_kwargs(s):
    p = s.pattern
    return dict(keyword=p.get("keyword"), wildcard=p.get("wildcard", 0), char_seq=p.get("char_seq", ()),
                values=p.get("values"))


# ------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the unmodified reference engine on the host cores
# ------------------------------------------------------------------------------------------------

def reference_step_seconds(w, blob_path, threads, block):
    """One step of the workload with mmoore::SearchEngine<T>::run (all searches); -> seconds, matches."""
    from _oracle import Ref
    total, matches = 0.0, 0
    for s in w.searches:
        r = Ref.engine(w.bits, blob_path, big_endian=s.big_endian, threads=threads, block=block, count_only=True,
                       **pattern_kwargs(s))
        total += r["seconds"]          # SearchEngine::run alone, not the conversion of its maps to Python objects
        matches += r["count"]
    return total, matches


def cpu_sample(w, total, cap_bytes):
    """Bounded sample of the workload for the CPU legs: the first min(total, cap) bytes of the file."""
    import monkey_moore_b200.workloads as wl
    n = min(total, cap_bytes)
    ws = w.scaled(n)
    blob = wl.host_blob(w, 0, n, total)          # a prefix of the very file the GPU arm scans
    d = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
    path = os.path.join(d, "mmoore_bench_%d_%s.bin" % (os.getpid(), w.key))
    blob.tofile(path)
    return ws, path, blob, "first %d MiB of the %s file, in %s, all searches of a step" % (n >> 20, w.key, d)


def cpu_baseline(w, total, cap_bytes=512 << 20, repeats=1):
    """The reference engine (all host cores) on a bounded sample, at the struct-default block and at the GUI's."""
    from _oracle import Ref
    if not Ref.available():
        return None
    cores = os.cpu_count() or 1
    ws, path, blob, sample = cpu_sample(w, total, cap_bytes)
    by_block = {}
    try:
        for block in sorted({w.block_size, GUI_BLOCK}):
            reference_step_seconds(ws, path, cores, block)          # page cache + first-touch warm-up
            ts = [reference_step_seconds(ws, path, cores, block)[0] for _ in range(repeats)]
            by_block[str(block)] = ws.size * len(ws.searches) / (sum(ts) / len(ts)) / 1e9
        one = None
        if w.key == "cfg1":       # the reference's own benchmark: MonkeyMoore<T>::search, one thread, in memory
            s = w.searches[0]
            data = blob if w.bits == 8 else blob[: len(blob) // 2 * 2].view(np.uint16)
            t, _m = Ref.time_search(w.bits, data, iters=3, **pattern_kwargs(s))
            one = ws.size / t / 1e9
    finally:
        os.unlink(path)
    out = {"value": by_block[str(w.block_size)], "unit": "GB/s", "cores": cores, "kind": "reference",
           "sample": sample + ", SearchEngine::run with %d threads" % cores, "by_block_size": by_block}
    if one is not None:
        out["search_1thread"] = one
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    import monkey_moore_b200.workloads as wl
    from _oracle import Ref
    if not Ref.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libmmref.so missing"}))
        return
    cores = os.cpu_count() or 1

    def arm(w, steps, warmup):
        scaling, total, _per = resolve_scaling(args, w, world)
        ws, path, _blob, sample = cpu_sample(w, total, 512 << 20)
        try:
            for _ in range(warmup):
                reference_step_seconds(ws, path, cores, w.block_size)
            times = [reference_step_seconds(ws, path, cores, w.block_size)[0] for _ in range(steps)]
            gui = reference_step_seconds(ws, path, cores, GUI_BLOCK)[0]
        finally:
            os.unlink(path)
        sec = sum(times) / len(times)
        nbytes = ws.size * len(ws.searches)
        return {"value": nbytes / sec / 1e9, "ms_per_step": sec * 1e3, "config": config_of(w, scaling, total, world),
                "dtype": dtype_of(w),
                "cpu_baseline": {"value": nbytes / sec / 1e9, "unit": "GB/s", "cores": cores, "kind": "reference",
                                 "sample": sample + ", SearchEngine::run with %d threads" % cores,
                                 "by_block_size": {str(w.block_size): nbytes / sec / 1e9, str(GUI_BLOCK): nbytes / gui / 1e9}}}

    w = wl.WORKLOADS[args.workload]
    head = arm(w, args.steps, args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": head["value"], "unit": "GB/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True,
            "scaling": head["config"]["scaling"], "vs_baseline": None, "dtype": head["dtype"], "data": "synthetic",
            "config": head["config"], "cpu_baseline": head["cpu_baseline"],
            "e2e": {"value": head["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    per = []
    for key in per_config_keys(args, world):
        r = arm(wl.WORKLOADS[key], 1, 1)
        per.append({"workload": key, "value": r["value"], "unit": "GB/s", "config": r["config"], "dtype": r["dtype"],
                    "cpu_baseline": r["cpu_baseline"]})
    if per:
        line["per_config"] = per
    print(json.dumps(line))


def per_config_keys(args, world):
    if args.per_config == "none":
        return []
    if args.per_config != "auto":
        return [k for k in args.per_config.split(",") if k]
    if args.workload != "cfg2" or args.size_mib:
        return []
    return ["cfg1", "cfg3", "cfg4", "cfg5"] if world == 1 else ["cfg4", "cfg5"]


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------

class Ctx:
    pass


def measure(ctx, args, w, steps, warmup, headline):
    """Everything for ONE workload on this job's GPUs -> dict (meaningful on rank 0)."""
    torch, mm, dist, comm = ctx.torch, ctx.mm, ctx.dist, ctx.comm
    rank, world, local, stream = ctx.rank, ctx.world, ctx.local, ctx.stream
    import monkey_moore_b200.workloads as wl
    from monkey_moore_b200.distributed import shard_bytes

    scaling, total_size, _per = resolve_scaling(args, w, world)
    W = w.bits // 8
    progs = [mm.Program(w.bits, **s.pattern) for s in w.searches]
    overlap = (max(p.keyword_len for p in progs) - 1) * W
    b0, nb, lo, hi = shard_bytes(total_size, w.block_size, overlap, rank, world)
    b1 = b0 + nb
    blob = wl.device_blob(w, first_byte=lo, nbytes=hi - lo, total_size=total_size)
    assert blob.data_ptr() % 16 == 0
    # inputs smaller than the L2 are rotated over distinct copies so that every scan streams from HBM
    copies = 1 if (hi - lo) >= 2 * L2_BYTES else -(-2 * L2_BYTES // max(hi - lo, 1))
    blobs = [blob] + [blob.clone() for _ in range(copies - 1)]
    turn = [0]
    mm.set_stream(stream.cuda_stream, True)
    torch.cuda.synchronize()

    def next_blob():
        turn[0] += 1
        return blobs[turn[0] % copies]

    def enqueue_step():
        # the searches of a step are independent: all of them are enqueued at once (no host wait in between)
        b = next_blob()
        return [prog.engine_scan(b, w.block_size, big_endian=s.big_endian, file_size=total_size,
                                 first_block=b0, num_blocks=nb, asynchronous=True)
                for prog, s in zip(progs, w.searches)]

    gather_ms = [0.0, 0]
    last_gather = [None]

    def complete_step(held, keep=None):
        """host side of a step: statistics, the gather (N > 1), release.  keep: list that receives what rank 0 got."""
        launches, filt_ms, filt_bytes, scan_ms = 0, 0.0, 0, 0.0
        for res in held:
            st = res.stats()
            launches += st["launches"]
            filt_ms += st["ms_filter"]
            scan_ms += st["ms_total"]
            filt_bytes += st["bytes_scanned"] + 8 * res.count
        if world > 1:
            g = comm.gather(held, lazy=True)       # ONE grouped NCCL op per step, on the communicator's own stream
            if keep is not None:
                keep.append(g)
            else:
                # rank 0 reads the headers of this gather when the NEXT one starts (or in comm.wait()): dropping the
                # previous object here costs nothing, dropping this one now would wait for the packed buffers
                old, last_gather[0] = last_gather[0], g
                if old is not None:
                    old.close()
        for r in held:
            r.close()
        return launches, filt_ms, filt_bytes, scan_ms

    def barrier():
        if world > 1:
            comm.wait()
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(warmup, 3)):
        complete_step(enqueue_step())
    barrier()
    sampler = None
    if headline:
        sampler = ClockSampler(local)
        sampler.start()
        time.sleep(0.15)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches, filt_ms, filt_bytes, scan_ms = 0, 0.0, 0, 0.0
    barrier()
    e0.record(stream)
    t0 = time.perf_counter()
    # software pipeline over the steps: the next step(s) are enqueued before the host collects step k (counts,
    # statistics, gather), so the device never waits for the host between steps; every step is complete before the
    # clock stops
    inflight = []
    depth = 2 if world > 1 else 1          # N > 1: the gather of step k travels while steps k+1 and k+2 scan
    for _ in range(steps):
        inflight.append(enqueue_step())
        if len(inflight) > depth:
            a = complete_step(inflight.pop(0))
            launches += a[0]; filt_ms += a[1]; filt_bytes += a[2]; scan_ms += a[3]
    while inflight:
        a = complete_step(inflight.pop(0))
        launches += a[0]; filt_ms += a[1]; filt_bytes += a[2]; scan_ms += a[3]
    e1.record(stream)
    if world > 1:
        gather_ms[0] = comm.wait()             # the last step's lists have landed on rank 0 / left this rank
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    dev_ms = e0.elapsed_time(e1)
    barrier()
    clocks = None
    if sampler is not None:
        sampler.stop_flag.set()
        sampler.join()
        clocks = sampler.summary()
    sec = max(wall, dev_ms / 1e3)
    if world > 1:
        t = torch.tensor([sec, gather_ms[0]], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sec, gmax = float(t[0].item()), float(t[1].item())
        lt = torch.tensor([launches], dtype=torch.int64, device="cuda")
        dist.all_reduce(lt)
        launches = int(lt.item())
    else:
        gmax = 0.0
    bytes_per_step = total_size * len(w.searches)
    value = bytes_per_step * steps / sec / 1e9

    # ---- roofline leg: the filter kernel ALONE.  Inside the timed region consecutive scans overlap on two streams
    # (the resolve of one beside the filter of the next), so the per-kernel CUDA events there measure kernels that
    # share the memory system.  Here every scan is completed before the next is enqueued; the events sit on the
    # stream the kernel is launched on, around the filter launch only.
    alone_ms, alone_bytes, alone_n, alone_total = 0.0, 0, 0, 0.0
    for _ in range(10 if (hi - lo) < (2 << 30) else 4):
        for prog, s_ in zip(progs, w.searches):
            res = prog.engine_scan(next_blob(), w.block_size, big_endian=s_.big_endian, file_size=total_size,
                                   first_block=b0, num_blocks=nb)
            st = res.stats()
            alone_ms += st["ms_filter"]
            alone_total += st["ms_total"]
            alone_bytes += st["bytes_scanned"] + 8 * res.count
            alone_n += 1
            res.close()
    torch.cuda.synchronize()
    # sparse scans resolve inside the filter kernel (one launch): the same kernel WITHOUT that tail, for comparison
    unf_ms, unf_bytes, unf_n, kinds = 0.0, 0, 0, set()
    old = mm.set_path_override(4)
    try:
        for _ in range(4):
            for prog, s_ in zip(progs, w.searches):
                res = prog.engine_scan(next_blob(), w.block_size, big_endian=s_.big_endian, file_size=total_size,
                                       first_block=b0, num_blocks=nb)
                st = res.stats()
                unf_ms += st["ms_filter"]; unf_bytes += st["bytes_scanned"] + 8 * res.count; unf_n += 1
                res.close()
    finally:
        mm.set_path_override(old)
    for prog, s_ in zip(progs, w.searches):
        res = prog.engine_scan(next_blob(), w.block_size, big_endian=s_.big_endian, file_size=total_size,
                               first_block=b0, num_blocks=nb)
        kinds.add(res.stats()["resolve_kind"])
        res.close()
    torch.cuda.synchronize()

    # ---- e2e: host (pinned) buffers through the public call, H2D + scan + D2H of the results (+ gather)
    e2e = None
    host_np = None
    pinned_total = (hi - lo) * min(world, torch.cuda.device_count())
    if not args.no_e2e and not fits_in_host_ram(pinned_total):
        e2e = {"value": None, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
               "note": "skipped: %d MiB of pinned host memory for this node's ranks would not leave enough free RAM" % (pinned_total >> 20)}
    elif not args.no_e2e:
        mm.set_stream(None, False)
        host = torch.empty(hi - lo, dtype=torch.uint8).pin_memory()
        host.copy_(blob)
        torch.cuda.synchronize()
        host_np = host.numpy()
        e2e_steps = 3 if (hi - lo) <= (1 << 30) else 2
        d2h = [0]

        def e2e_step():
            d2h[0] = 0
            held = []
            for prog, s in zip(progs, w.searches):
                res = prog.engine_scan(host_np, w.block_size, big_endian=s.big_endian, file_size=total_size,
                                       first_block=b0, num_blocks=nb)
                off, val = res.arrays()
                d2h[0] += off.nbytes + val.nbytes // 2
                held.append(res)
            if world > 1:
                g = comm.gather(held, lazy=True)
                if g is not None:
                    g.counts()
                    g.close()
                comm.wait()
            for r in held:
                r.close()
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        if world > 1:
            comm.wait()
        torch.cuda.synchronize()
        e2e_sec = time.perf_counter() - t0
        barrier()
        if world > 1:
            t = torch.tensor([e2e_sec, float(d2h[0])], dtype=torch.float64, device="cuda")
            dist.all_reduce(t[:1], op=dist.ReduceOp.MAX)
            dist.all_reduce(t[1:], op=dist.ReduceOp.SUM)
            e2e_sec, d2h_all = float(t[0].item()), int(t[1].item())
        else:
            d2h_all = d2h[0]
        e2e = {"value": bytes_per_step * e2e_steps / e2e_sec / 1e9, "unit": "GB/s",
               "h2d_bytes_per_step": int(bytes_per_step + (world - 1) * overlap * len(w.searches)),
               "d2h_bytes_per_step": int(d2h_all), "steps": e2e_steps,
               "note": "pinned host buffer -> public engine_scan call per search (each search copies the slice H2D, like "
                       "the reference re-reads the file per search) -> result arrays on the host"}
        mm.set_stream(stream.cuda_stream, True)

    # ---- parity at full size against the CPU oracle
    parity = None
    if not args.no_verify:
        parity = verify(ctx, w, progs, blob, host_np, total_size, b0, nb, lo, hi, complete_step, enqueue_step)

    # ---- CPU baseline beside it (rank 0 only): the reference engine on a bounded sample
    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        try:
            cpu = cpu_baseline(w, total_size)
        except Exception as e:  # the baseline is reported, never required
            cpu = {"value": None, "unit": "GB/s", "cores": os.cpu_count(), "kind": "reference", "sample": "failed: %s" % e}

    mm.set_stream(None, False)
    del blobs, blob
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    peak, peak_src = measured_peak()
    n_filter = steps * len(w.searches)
    achieved = (alone_bytes / alone_n) / (alone_ms / alone_n) / 1e6          # GB/s, per-launch averages, kernel alone
    achieved_overlapped = (filt_bytes / n_filter) / (filt_ms / n_filter) / 1e6   # same events inside the timed region
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f).get(w.key)
    except Exception:
        pass
    l2 = ("slice (%d MiB per GPU) larger than the 126 MB L2, no flush needed" % ((hi - lo) >> 20) if copies == 1 else
          "slice of %d MiB rotated over %d distinct device copies (%d MiB > 2 x L2), so every scan streams from HBM"
          % ((hi - lo) >> 20, copies, (copies * (hi - lo)) >> 20))
    return {"value": value, "ms_per_step": sec / steps * 1e3, "device_ms_per_step": dev_ms / steps,
            "steps": steps, "warmup": max(warmup, 3), "scaling": scaling, "dtype": dtype_of(w),
            "config": config_of(w, scaling, total_size, world), "l2": l2, "parallelism": "block-sharded x%d" % world,
            "scan_ms_per_step": scan_ms / steps, "gather_ms_last_step": gmax,
            "scan_alone_gbs": (alone_bytes / alone_n) / (alone_total / alone_n) / 1e6,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "frac_of_nominal_8TBs": achieved / 8000.0, "traffic": traffic,
                         "kernel": "k_filter" if w.bits == 16 else "k_filter8", "peak_source": peak_src,
                         "timing": "CUDA events around the filter launch on its stream, %d launches run alone after the "
                                   "timed region (inside it scans overlap on two streams)" % alone_n,
                         "achieved_in_timed_region": achieved_overlapped,
                         "resolve_fused_into_this_kernel": kinds == {1},
                         "filter_without_fused_resolve": (unf_bytes / unf_n) / (unf_ms / unf_n) / 1e6,
                         "whole_scan_frac": (alone_bytes / alone_n) / (alone_total / alone_n) / 1e6 / peak,
                         "pipeline_frac": value / world / peak},
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "parity": parity}


def verify(ctx, w, progs, blob, host_np, total_size, b0, nb, lo, hi, complete_step, enqueue_step):
    """Full-size parity: this rank's lists against the oracle on the same blocks; rank 0 also checks the gathered
    lists against the rank-order composition of every rank's oracle digest."""
    torch, dist, rank, world = ctx.torch, ctx.dist, ctx.rank, ctx.world
    import monkey_moore_b200.workloads as wl
    from _oracle import Oracle, compose, digest
    threads = max(1, (os.cpu_count() or 1) // world)
    patches = wl.planted_patches(w, total_size) if w.generator == "splitmix" else None
    held = enqueue_step()
    mine, want, ok_local, full_lists = [], [], True, 0
    for prog, s, res in zip(progs, w.searches, held):
        off, val = res.arrays()
        d = digest(off, val)
        o = Oracle(w.bits, **pattern_kwargs(s))
        if patches is None:        # host-generated blob (cfg1): the in-memory oracle engine on the same bytes
            hb = host_np if host_np is not None else blob.cpu().numpy()
            assert world == 1
            eo, ev = o.engine(hb, w.block_size, big_endian=s.big_endian, wrap32=False)
            e = digest(eo, ev)
            ok_local = ok_local and off.tolist() == eo.tolist() and val.tolist() == ev.tolist()
            full_lists += 1
        else:
            small = d[0] <= 10_000_000
            r = o.engine_synth(w.seed, w.byte_mask, total_size, w.block_size, b0, nb, s.big_endian, patches,
                               threads=threads, want_list=small, list_cap=max(d[0], 1) + 16)
            e = (r["count"], r["s0"], r["s1"])
            if small and r["count"] == d[0]:
                ok_local = ok_local and off.tolist() == r["offsets"].tolist() and val.tolist() == r["values"].tolist()
                full_lists += 1
        ok_local = ok_local and d == e
        mine.append(d)
        want.append(e)
    gathered_ok, total_matches = None, sum(d[0] for d in mine)
    if world > 1:
        keep = []
        complete_step(held, keep)
        # every rank's oracle digests -> rank 0
        t = torch.tensor([[int(x) - (1 << 64) if int(x) >= (1 << 63) else int(x) for x in e] for e in want],
                         dtype=torch.int64, device="cuda")
        allw = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allw, t)
        okt = torch.tensor([1 if ok_local else 0], dtype=torch.int64, device="cuda")
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
        ok_local = bool(okt.item())
        if rank == 0:
            g = keep[0]
            got = g.fetch()
            gathered_ok, total_matches = True, 0
            for k, (off, val) in enumerate(got):
                parts = [tuple(int(x) & ((1 << 64) - 1) for x in allw[r][k].tolist()) for r in range(world)]
                gathered_ok = gathered_ok and digest(off, val) == compose(parts)
                gathered_ok = gathered_ok and bool(np.all(off[1:] > off[:-1]))
                total_matches += len(off)
            g.close()
        ctx.comm.wait()
    else:
        for r in held:
            r.close()
    if rank != 0:
        return None
    return {"checked": ("every rank: its lists vs the CPU oracle on the same blocks at full size (count + order-sensitive "
                        "digest%s)%s" % (", full lists where <= 1e7 matches" if full_lists else "",
                                         "; rank 0: the gathered lists vs the rank-order composition of all ranks' oracle "
                                         "digests, offsets strictly ascending" if world > 1 else "")),
            "bit_exact": bool(ok_local and (gathered_ok is None or gathered_ok)), "per_rank_ok": bool(ok_local),
            "gathered_ok": gathered_ok, "matches": int(total_matches),
            "digest": ["%016x" % d[2] for d in mine]}


def measure_chain(ctx, args):
    """the cfg4 bytes searched as ONE MonkeyMoore<uint16_t>::search chain (even alignment, LE) instead of
    independent engine blocks.  N = 1: the whole buffer in one call, and in four slices through mmg_chain_*;
    N > 1: one slice per rank through mmg_comm_search (the slice maps cross the ranks in one 128-byte all-gather),
    lists gathered to rank 0.  Parity: every rank's list against the oracle's chain over the same slice entered with
    the same phase, the oracle's exit phase of rank r against the entry phase rank r + 1 was given, entry 0 = 0."""
    torch, mm, dist, comm = ctx.torch, ctx.mm, ctx.dist, ctx.comm
    rank, world = ctx.rank, ctx.world
    import monkey_moore_b200.workloads as wl
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from _oracle import Oracle, digest

    w = wl.WORKLOADS["cfg4"]
    s = w.searches[0]
    total = (w.single_gpu_size or w.size) if world == 1 else w.size
    if args.size_mib:
        total = args.size_mib << 20
    W = w.bits // 8
    prog = mm.Program(w.bits, **s.pattern)
    n = total // W
    tail = prog.keyword_len - 1
    per = (n // world) * W // 4096 * 4096 // W
    first = rank * per
    owned = per if rank < world - 1 else n - first
    avail = owned if rank == world - 1 else owned + tail
    blob = wl.device_blob(w, first_byte=first * W, nbytes=avail * W, total_size=total)
    mm.set_stream(ctx.stream.cuda_stream, True)

    def barrier():
        if world > 1:
            comm.wait()
            dist.barrier()
        torch.cuda.synchronize()

    def step(keep=None):
        if world > 1:
            res = comm.search(prog, blob, owned, first)
            g = comm.gather([res], lazy=True)
            if keep is not None:
                keep.append((res, g))
                return
            res.close()
            if g is not None:
                g.close()
        else:
            res = prog.search(blob)
            if keep is not None:
                keep.append((res, None))
                return
            res.count
            res.close()

    def timed(fn, steps):
        for _ in range(3):
            fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        barrier()
        sec = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([sec], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            sec = float(t.item())
        return sec

    steps = 5
    sec = timed(step, steps)
    out = {"workload": "cfg4 bytes as ONE MonkeyMoore<uint16_t>::search chain (even alignment)", "bytes": total,
           "slices": world, "steps": steps, "value": round(total * steps / sec / 1e9, 1), "unit": "GB/s",
           "ms_per_step": round(sec / steps * 1e3, 3),
           "timing": "host clock around the steps, barrier + synchronize on both sides, max over ranks (every step waits "
                     "for the map exchange on the host)" if world > 1 else "host clock, synchronize on both sides"}
    if world == 1:
        def sliced():
            parts = prog.search_sliced(blob, (n // 4) * W // 4096 * 4096 // W)
            for p_ in parts:
                p_.count
                p_.close()
        sec4 = timed(sliced, steps)
        out["four_slices_one_gpu"] = {"value": round(total * steps / sec4 / 1e9, 1), "ms_per_step": round(sec4 / steps * 1e3, 3)}
    if args.no_verify:
        return out
    # ---- parity
    keep = []
    step(keep)
    res, g = keep[0]
    off, val = res.arrays()
    entry = res.stats()["chain_entry"]
    if not fits_in_host_ram(avail * W * (min(world, torch.cuda.device_count()))):
        out["parity"] = {"checked": False, "note": "host copy of the slices would not fit"}
        return out
    host = blob.cpu().numpy().view(np.uint16)
    o = Oracle(w.bits, keyword=s.pattern.get("keyword"), wildcard=s.pattern.get("wildcard", 0), char_seq=s.pattern.get("char_seq", ()))
    t0 = time.perf_counter()
    opos, oval, oexit = o.search_slice(host, entry, owned)
    cpu_sec = time.perf_counter() - t0
    same = off.tolist() == (opos + np.uint64(first)).tolist() and val.tolist() == oval.tolist()
    ok = {"rank_lists_equal_oracle": bool(same), "matches": int(len(off))}
    if world > 1:
        t = torch.tensor([entry, oexit - owned if rank < world - 1 else 0, int(same), len(off)], dtype=torch.int64, device="cuda")
        parts = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(parts, t)
        rows = [p_.tolist() for p_ in parts]
        linked = rows[0][0] == 0 and all(rows[r + 1][0] == rows[r][1] for r in range(world - 1))
        ok = {"rank_lists_equal_oracle": all(r[2] == 1 for r in rows), "entry_phases": [r[0] for r in rows],
              "entry_phases_equal_oracle_exits": bool(linked), "matches": sum(r[3] for r in rows)}
        if rank == 0:
            goff, _gval = g.fetch()[0]
            ok["gathered_count"] = int(len(goff))
            ok["gathered_ascending"] = bool(np.all(goff[1:] > goff[:-1])) if len(goff) > 1 else True
            ok["gathered_equals_sum_of_ranks"] = int(len(goff)) == ok["matches"]
            ok["gathered_prefix_equals_rank0_list"] = goff[: len(off)].tolist() == off.tolist()
            g.close()
    else:
        parts = prog.search_sliced(blob, (n // 4) * W // 4096 * 4096 // W)
        soff = np.concatenate([p_.arrays()[0] for p_ in parts])
        ok["four_slices_equal_whole"] = soff.tolist() == off.tolist()
        ok["entry_phases"] = [p_.stats()["chain_entry"] for p_ in parts]
        for p_ in parts:
            p_.close()
    res.close()
    ok["oracle_cpu_gbs_1thread"] = round(owned * W / cpu_sec / 1e9, 2)
    out["parity"] = ok
    barrier()
    return out


def run_ours(args):
    import torch

    import monkey_moore_b200 as mm
    import monkey_moore_b200.workloads as wl

    ctx = Ctx()
    ctx.torch, ctx.mm = torch, mm
    ctx.rank = int(os.environ.get("RANK", "0"))
    ctx.world = int(os.environ.get("WORLD_SIZE", "1"))
    ctx.local = int(os.environ.get("LOCAL_RANK", "0"))
    if mm.device_count() == 0:
        raise SystemExit("bench.py: no CUDA device and no CPU fallback exists for the product path")
    torch.cuda.set_device(ctx.local)
    ctx.dist, ctx.comm = None, None
    if ctx.world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", ctx.local))
        ctx.dist = dist

        def bcast(raw):
            t = torch.tensor(list(raw), dtype=torch.uint8, device="cuda")
            dist.broadcast(t, src=0)
            return bytes(t.cpu().tolist())
        ctx.comm = mm.Comm(ctx.rank, ctx.world, bcast)     # the library's own NCCL communicator for the result gather
    ctx.stream = torch.cuda.current_stream()

    w = wl.WORKLOADS[args.workload]
    head = measure(ctx, args, w, args.steps, args.warmup, True)
    per = []
    for key in per_config_keys(args, ctx.world):
        wk = wl.WORKLOADS[key]
        _sc, tot, _p = resolve_scaling(args, wk, ctx.world)
        ksteps = 50 if tot <= (64 << 20) else (10 if tot // ctx.world <= (1 << 30) else 5)
        try:
            r = measure(ctx, args, wk, ksteps, 3, False)
        except Exception as e:          # a failing side configuration must not cost the headline line
            r = {"error": "%s: %s" % (type(e).__name__, e)} if ctx.rank == 0 else None
            if ctx.world > 1:
                raise
        if r is not None:
            r["workload"] = key
            r["unit"] = "GB/s"
            per.append(r)
    chain = None
    if args.chain == "on" or (args.chain == "auto" and per_config_keys(args, ctx.world)):   # auto: with the default full run
        try:
            chain = measure_chain(ctx, args)
        except Exception as e:
            if ctx.world > 1:
                raise
            chain = {"error": "%s: %s" % (type(e).__name__, e)}
    if ctx.rank == 0:
        line = {"metric": METRIC, "value": head["value"], "unit": "GB/s", "n_gpus": ctx.world, "steps": head["steps"],
                "warmup": head["warmup"], "ms_per_step": head["ms_per_step"], "higher_is_better": True,
                "scaling": head["scaling"], "vs_baseline": None, "dtype": head["dtype"], "data": "synthetic",
                "config": head["config"]}
        for k in ("l2", "parallelism", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks", "parity",
                  "device_ms_per_step", "scan_ms_per_step", "gather_ms_last_step", "scan_alone_gbs"):
            line[k] = head[k]
        if per:
            line["per_config"] = per
        if chain is not None:
            line["single_chain"] = chain
        print(json.dumps(line))
    if ctx.world > 1:
        ctx.comm.close()
        ctx.dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
