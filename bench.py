#!/usr/bin/env python3
"""bench.py -- throughput of the annotate hot path (BASELINE.json metric: reads/s and Gbases/s annotated).

A "step" is one pass of the whole hot path (flank scan -> local minima -> traceback -> barcode stage -> collapse) over
one batch of synthetic reads of configs[1] (SQK-NBD114-96, 10 kb reads).  `value` is measured with the batch resident
in HBM (CUDA events on the launching stream); `e2e` goes through the pipelined C-ABI calls with pinned HOST buffers
(host->device copy of the reads and device->host copy of the rows inside the timed region).

  python bench.py --gpus 1 --steps 5 --warmup 3
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...      (one rank per GPU, weak scaling)
  python bench.py --impl reference ...     times the CPU restatement (oracle) on the host cores -- the reference is
                                           Rust + crates.io dependencies and cannot be built in this image.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np

KIT = "SQK-NBD114-96"
READ_LEN = 10000
WORKLOAD = "SQK-NBD114-96 (96 native barcodes + flanks), synthetic 10 kb reads, 1xB200 per rank (BASELINE configs[1])"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.gpu, self.lines, self.proc = gpu, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def mark(self):
        return time.time()

    def stop(self, t0=None, t1=None):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        lines = [ln for ts, ln in self.lines if t0 is None or (t0 <= ts <= t1 + 0.05)]
        window = "timed region"
        if len(lines) < 3:                      # region shorter than the sampling period: use warm-up + timed region (same load)
            lines = [ln for ts, ln in self.lines]
            window = "warm-up + timed region"
        for ln in lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx = float(f[2])
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm), window=window)


def make_batch(groups, n_reads, seed):
    from barbell_b200 import synth
    return synth.make_reads(groups, n_reads, READ_LEN, seed=seed)


def run_reference(args):
    """The reference arm: the CPU restatement of the same path on all host cores, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import barbell_b200 as bb
    import oracle_lib as O
    from barbell_b200 import synth
    gs = bb.GroupSet.from_kit(KIT)
    G = gs.as_dicts()
    threads = O.lib().orc_max_threads()
    n = args.ref_reads
    bases, offsets, _ = make_batch(G, n, synth.SEED0 + 2)
    for _ in range(max(1, min(args.warmup, 1))):
        O.demux_batch(G, bases[:int(offsets[min(n, 64)])], offsets[:min(n, 64) + 1], n_threads=threads)
    t0 = time.perf_counter()
    rows = 0
    for _ in range(args.steps):
        rows += len(O.demux_batch(G, bases, offsets, n_threads=threads))
    dt = time.perf_counter() - t0
    v = n * args.steps / dt
    sample = f"{n} synthetic 10 kb reads per step x {args.steps} steps, oracle (C restatement, bit-vector scan), {threads} threads"
    out = dict(metric="reads_per_s", value=v, unit="reads/s", impl="reference", n_gpus=args.gpus, steps=args.steps,
               warmup=args.warmup, ms_per_step=1e3 * dt / args.steps, higher_is_better=True, scaling="weak",
               vs_baseline=None, dtype="u64 bit-vectors (integer) + f64 score", data="synthetic",
               gbases_per_s=v * READ_LEN / 1e9,
               config=dict(workload=WORKLOAD, reads_per_step=n, read_len=READ_LEN, kit=KIT),
               cpu_baseline=dict(value=v, unit="reads/s", cores=threads, kind="port", sample=sample),
               e2e=dict(value=v, unit="reads/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0), rows=rows)
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--reads", type=int, default=100000, help="reads per device-resident batch (10 kb each: 1 GB)")
    ap.add_argument("--ref-reads", type=int, default=20000, help="reads per step of the CPU arm (--impl reference)")
    ap.add_argument("--cpu-reads", type=int, default=40000, help="reads of the batch timed on the host cores for cpu_baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-sub", type=int, default=4, help="sub-batches per step of the end-to-end leg")
    ap.add_argument("--e2e-depth", type=int, default=4, help="sub-batches in flight (<= BB_MAX_INFLIGHT = 4)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    import barbell_b200 as bb
    from barbell_b200 import sharding, synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: barbell_b200 has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        # the ranks share the host cores: split them for the host-side packing of the end-to-end leg
        os.environ.setdefault("BB_PACK_THREADS", str(max(1, (os.cpu_count() or 1) // int(os.environ.get("LOCAL_WORLD_SIZE", world)))))

    gs = bb.GroupSet.from_kit(KIT)
    G = gs.as_dicts()
    an = bb.Annotator(gs, device=local)
    n_reads = args.reads
    bases, offsets, _ = make_batch(G, n_reads, synth.SEED0 + 2 + 1000 * rank)
    total = int(offsets[-1])
    algo_bytes = total + 8 * n_reads                       # L + 8 bytes per read
    stream = torch.cuda.current_stream()
    d_bases = torch.from_numpy(bases).cuda()
    d_offsets = torch.from_numpy(offsets.astype(np.int64)).cuda()

    def step():
        return an.annotate_device(d_bases.data_ptr(), d_offsets.data_ptr(), n_reads, total, stream.cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(args.warmup):
        n_rows = step()
    barrier()
    t_mark0 = sampler.mark()
    l0 = an.kernel_launches()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    scan_ms = []
    stage_acc = {}
    ev0.record(stream)
    for _ in range(args.steps):
        n_rows = step()
        st = an.stage_ms()
        scan_ms.append(st["scan"])
        for k, v in st.items():
            stage_acc[k] = stage_acc.get(k, 0.0) + v
    ev1.record(stream)
    barrier()
    clocks = sampler.stop(t_mark0, sampler.mark())
    launches = an.kernel_launches() - l0
    ms = ev0.elapsed_time(ev1)
    hist_local = sharding.label_histogram(an.fetch_rows(int(n_rows)), G)     # rows of the last timed step (outside the timed region)
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * n_reads * args.steps / (ms_max / 1e3)

    # ---- end to end: pinned host buffers through bb_submit / bb_collect (4 sub-batches in flight), copies inside the timed region ----
    e2e = None
    if not args.no_e2e:
        n_sub, depth = args.e2e_sub, args.e2e_depth
        cuts = np.linspace(0, n_reads, n_sub + 1).astype(int)
        subs = []
        for i in range(n_sub):
            lo, hi = int(cuts[i]), int(cuts[i + 1])
            hb = torch.from_numpy(bases[int(offsets[lo]):int(offsets[hi])]).pin_memory()
            ho = torch.from_numpy((offsets[lo:hi + 1] - offsets[lo]).astype(np.int64)).pin_memory()
            subs.append((hb, ho, hi - lo))
        h2d = sum(hb.numel() + ho.numel() * 8 for hb, ho, _ in subs)

        def e2e_pass(n_steps):
            rows = 0
            jobs = [(s, i) for s in range(n_steps) for i in range(n_sub)]
            inflight = nxt = 0
            while nxt < len(jobs) or inflight:
                while nxt < len(jobs) and inflight < depth:
                    hb, ho, nr = subs[jobs[nxt][1]]
                    an.submit(hb.data_ptr(), ho.data_ptr(), nr, tag=nxt)
                    nxt += 1; inflight += 1
                _, _, n = an.collect(copy=False)
                rows += n; inflight -= 1
            return rows

        def timed_e2e(annot):
            nonlocal an
            an_saved, an = an, annot
            try:
                e2e_pass(2)                              # warm-up (the packed mode also settles its head/tail split here)
                barrier()
                b0 = annot.h2d_bytes()
                t0 = time.perf_counter()
                rows = e2e_pass(args.steps)
                torch.cuda.synchronize()
                dt = time.perf_counter() - t0
                moved = (annot.h2d_bytes() - b0) // args.steps
            finally:
                an = an_saved
            te = torch.tensor([dt], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(te, op=dist.ReduceOp.MAX)
            return float(te.item()), rows, int(moved)

        dt_plain, rows_e2e, moved_plain = timed_e2e(an)
        # same call with bb_opts.flags bit 1: the library nibble-packs the HEAD of every batch on the host cores while the tail is
        # copied as it is (the split adapts to the measured pack and link rates), so PCIe moves fewer bytes
        an_pack = bb.Annotator(gs, device=local, pack_h2d=True)
        dt_pack, rows_pack, moved_pack = timed_e2e(an_pack)
        an_pack.close()
        # flags bit 2: the denser wire format, 2 bits per base for A/C/G/T + an exception list for every other byte
        an_crumb = bb.Annotator(gs, device=local, pack_h2d="crumbs")
        dt_crumb, rows_crumb, moved_crumb = timed_e2e(an_crumb)
        an_crumb.close()
        assert rows_pack == rows_e2e and rows_crumb == rows_e2e and moved_plain == h2d
        modes = {"plain": (dt_plain, moved_plain), "packed": (dt_pack, moved_pack), "crumbs": (dt_crumb, moved_crumb)}
        best = min(modes, key=lambda k: modes[k][0])
        dt = modes[best][0]
        e2e = dict(value=world * n_reads * args.steps / dt, unit="reads/s", h2d_bytes_per_step=modes[best][1],
                   d2h_bytes_per_step=int(rows_e2e // args.steps * 88),
                   api=f"bb_submit/bb_collect, pinned host buffers, {n_sub} sub-batches per step, {depth} in flight; mode=" + best +
                       (" (head of every batch packed by the library on the host cores -- " +
                        ("4 bits per base" if best == "packed" else "2 bits per base + exception list") +
                        " -- and expanded on the device, tail copied as it is; bytes counted by the library: bb_h2d_bytes)" if best != "plain" else ""),
                   gbases_per_s=world * total * args.steps / dt / 1e9,
                   by_mode={k: world * n_reads * args.steps / v[0] for k, v in modes.items()},
                   h2d_bytes_by_mode={k: v[1] for k, v in modes.items()})

    counters = an.counters()
    summed = sharding.all_reduce_counters(counters["total"], counters["kept"]) if world > 1 else counters
    # per-barcode row counts of the last step, summed over the ranks (the one other collective of the path)
    hist = sharding.all_reduce_label_counts(hist_local)

    if rank == 0:
        peak, peak_src = peaks()
        scan_avg = float(np.mean(scan_ms))
        achieved = algo_bytes / (scan_avg / 1e3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:    # dram__bytes_read+write per launch from the committed ncu --set full capture, scaled to this batch size
                traffic = json.load(open(tp))["k_flank_scan_dram_bytes_per_algorithmic_byte"] * algo_bytes
            except Exception:
                traffic = None
        out = dict(metric="reads_per_s", value=value, unit="reads/s", n_gpus=world, steps=args.steps, warmup=args.warmup,
                   ms_per_step=ms_max / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
                   dtype="u64 bit-vectors (integer) + f64 score", data="synthetic",
                   gbases_per_s=value * READ_LEN / 1e9,
                   config=dict(workload=WORKLOAD, kit=KIT, reads_per_step=n_reads, read_len=READ_LEN,
                               batch_bytes=total, l2_policy="batch (1 GB) larger than the 126 MB L2; same batch every step",
                               parallelism=f"reads sharded over {world} GPU(s), no data-path collective"),
                   roofline=dict(bound="hbm", achieved=achieved, peak=peak, unit="GB/s", frac=achieved / peak, traffic=traffic,
                                 kernel="flank scan stage (the kernels that stream the bases): k_flank_filter [scan + candidate runs + pre-check] + k_flank_verify (+ chunk index)",
                                 whole_step=dict(achieved=algo_bytes / (ms_max / args.steps / 1e3) / 1e9, frac=algo_bytes / (ms_max / args.steps / 1e3) / 1e9 / peak,
                                                 note="same algorithmic bytes over the WHOLE step; the barcode stage (k_barcode, ~60 % of the step) touches < 1 % of the bytes: "
                                                      "one warp per flank match aligns all 96 barcodes with traceback + Lodhi score, issue/ALU bound"),
                                 algorithmic_bytes_per_launch=algo_bytes, ms_per_launch=scan_avg, peak_source=peak_src,
                                 note="integer-issue bound (bit-vector DP, 23 instructions per base for both strands, ALU pipe 77 % active: "
                                      "profiles/r1_ncu_full_final_kernels.txt), not DRAM bound; see DESIGN.md section 3"),
                   stage_ms_per_step={k: v / args.steps for k, v in stage_acc.items()},
                   e2e=e2e, gpu_launches=int(launches), clocks=clocks, rows_per_step=int(n_rows), counters=summed,
                   label_counts=dict(labels_seen=int((hist[1:] > 0).sum()), rows=int(hist.sum()), flank_only_rows=int(hist[0]),
                                     note="rows per barcode of the last step, all ranks (all_reduce of %d int64)" % len(hist)))
        if not args.no_cpu_baseline:
            import oracle_lib as O
            threads = O.lib().orc_max_threads()
            nb = min(args.cpu_reads, n_reads)
            sb, so = sharding.slice_reads(bases, offsets, 0, nb)
            t0 = time.perf_counter()
            rows_cpu = O.demux_batch(G, sb, so, n_threads=threads)
            dt = time.perf_counter() - t0
            out["cpu_baseline"] = dict(value=nb / dt, unit="reads/s", cores=threads, kind="port",
                                       sample=f"first {nb} reads of the same batch, oracle (C restatement, bit-vector scan), {threads} threads, "
                                              f"{dt:.2f} s wall = {dt * threads:.0f} core-seconds")
            # the same reads through the GPU path must give the same rows
            chk = an.annotate(sb, so)
            out["cpu_baseline"]["rows_match_gpu"] = bool(chk.tobytes() == rows_cpu.tobytes())
        print(json.dumps(out), flush=True)
    an.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
