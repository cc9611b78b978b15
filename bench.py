#!/usr/bin/env python
"""bench.py -- BURST alignment hot path on B200: reads/s and DP GCUPS (BASELINE.json metric).

A "step" is one pass of the hot path (prefix filter -> banded extend/rescore -> select) over one
batch: BASELINE.json configs[1] shape, 1 M synthetic 100 bp reads with exactly 2 edits (the
reference simulator's model, embalmlets/LLsim.c) against a 2 GB synthetic .edx-layout database,
-i 0.98 (budget 2), BEST-style minimum selection, forward + reverse-complement strands, task list
= what the reference's accelerated driver enumerates (bunches of 16 sorted strands x the bunch's
candidate clumps, burst.c:4077-4157).

  value     whole-job reads/s with queries, tasks and DB resident in HBM (kernels only: k_init_best, k_seedw, k_bin_count/offsets/scatter,
            k_extend x 9 band classes, k_select = 15 launches per step)
  e2e       the same through bg_align_bunches_into(): pinned host buffers in (reads once, 2 bits per base), hits in pinned host memory out
  roofline  dominant kernel (k_seedw) algorithmic bytes / its CUDA-event time vs measured HBM peak
            -- the kernel is integer-ALU bound, see "alu" and DESIGN.md
  cpu_baseline / --impl reference
            the reference's own aded_mat16L + reScoreM_mat16 (oracle/_ref/libburstref.so, built
            from /root/reference/burst.c) driven in the reference's loop shape on the host cores

Launch: python bench.py [--gpus N --steps K --warmup W]   (N>1 under torch.distributed.run)
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads", type=int, default=1_000_000, help="reads per GPU per step")
    ap.add_argument("--db-mb", type=int, default=2048)
    ap.add_argument("--read-len", type=int, default=100)
    ap.add_argument("--edits", type=int, default=2)
    ap.add_argument("--clump-len", type=int, default=214)
    ap.add_argument("--cpu-sample-bunches", type=int, default=0, help="0 = auto (about 10-30 s of CPU work)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--inflight", type=int, default=3, help="contexts (batches in flight) per GPU in the e2e leg")
    ap.add_argument("--config", default="c2", choices=["c2", "target", "c3"],
                    help="c2 (default, the headline): BASELINE.json configs[1], 1 M reads vs a 2 GB DB; target: north_star's 10 M x 100 bp vs a 31.5 GB .edx on one B200; "
                         "c3: configs[2] shape, 200 k x 292 bp amplicon reads (0-5 substitutions, budget 9) vs a 70 MB mutation-tree DB of 1400-base references, ~60 clump visits per strand")
    a = ap.parse_args()
    a.budget = a.edits; a.probe_bunches = 256
    if a.config == "target":
        a.reads, a.db_mb = 10_000_000, 32256
    if a.config == "c3":
        a.reads, a.db_mb, a.read_len, a.edits, a.clump_len, a.budget, a.probe_bunches = 200_000, 70, 292, 5, 1400, 9, 16
    return a


PARITY_SAMPLE_MOD = 64        # the CPU reference also records the kept lanes of every 64th read, for the bench-size parity check
DPX_PEAK = 571.6e9 * 32      # VIADDMNMX.U32 thread-instructions/s, measured (profiles/r1d_pipe_microbench.txt)


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"], "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        self.rows = []
        self.index = index
        self.proc = None

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def bind_to_gpu_numa_node(torch, local):
    """Best effort: run this rank on the CPUs of the NUMA node its GPU hangs off, so that first-touch places the pinned
    host buffers there and N ranks do not funnel their host->device copies through one socket.  Returns the node or None."""
    try:
        p = torch.cuda.get_device_properties(local)
        addr = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % addr).read())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


def build_workload(args, rank):
    from burst_b200 import synth
    t0 = time.time()
    if args.config == "c3":
        w = synth.amplicon_workload(args.reads, args.read_len, args.edits, args.db_mb << 20, args.clump_len, seed=20261017 + rank, budget=args.budget, halo=7)
    else:
        w = synth.bunch_workload(args.reads, args.read_len, args.edits, args.db_mb << 20, args.clump_len, seed=20261017 + rank)
    w["gen_s"] = time.time() - t0
    return w


def cpu_reference(args, w, steps=1, warmup=0):
    """The reference's kernels on the host cores over a bounded sample of the same task list."""
    from oracle import pyoracle
    threads = os.cpu_count() or 1
    refout = None
    qb = w["qbunch"]
    nb_all = (len(w["qoff"]) - 1 + qb - 1) // qb
    if pyoracle.Reference.available():
        ref = pyoracle.Reference()
        kind = "reference"
        nb = args.cpu_sample_bunches
        if not nb:
            probe = min(nb_all, args.probe_bunches * threads)
            t0 = time.perf_counter(); pyoracle.reference_run_bunches(ref, w, probe, threads); dt = time.perf_counter() - t0
            nb = int(min(nb_all, max(probe, probe * 4.0 / max(dt, 1e-3))))     # ~4 s per step
        times = []
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            r = pyoracle.reference_run_bunches(ref, w, nb, threads, sample_mod=PARITY_SAMPLE_MOD)
            if it >= warmup:
                times.append(time.perf_counter() - t0)
        nq = r["nq"]
        refout = r
        found = int((r["ed"][np.unique(w["slot"][:nq])] <= args.budget).sum())
        desc = "reference kernels aded_mat16L+reScoreM_mat16 (burst.c) in the reference's bunch loop, first %d of %d bunches = %d strands (%d pass-1 calls, %d truncated, %d pass-2), %d threads" % (
            nb, nb_all, nq, r["calls"], r["truncated"], r["rescore"], threads)
    else:
        # port: scalar oracle, one thread per OpenMP worker, far smaller sample
        orc = pyoracle.Oracle()
        kind = "port"
        nb = args.cpu_sample_bunches or 8
        nq = min(len(w["qoff"]) - 1, nb * qb)
        tk = w["tasks"][w["tasks"][:, 0] < nq]
        times = []
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            orc.run_tasks(w["packed"], w["clump_off"], w["clump_len"], w["qcodes"], w["qoff"], w["budget"], w["slot"], w["nslots"],
                          tk[:, 0], tk[:, 1], orc.score_table(1), 0)
            if it >= warmup:
                times.append(time.perf_counter() - t0)
        desc = "scalar oracle port, first %d bunches = %d strands, %d tasks" % (nb, nq, len(tk))
    dt = float(np.mean(times))
    reads = nq / 2.0
    return {"value": reads / dt, "unit": "reads/s", "cores": threads, "kind": kind, "sample": desc,
            "seconds_per_step": dt, "reads_in_sample": reads}, dt, refout


def parity_vs_reference(w, runs, hits, best, refout):
    """The GPU's results against what the reference's own kernels produced on the same bunches (bench-size parity):
    per-slot minima of every slot whose strands were all inside the CPU sample, and -- for the sampled slots (slot % PARITY_SAMPLE_MOD
    == 0) -- the kept lanes with their (ed, numGapQ, numGapR, finalPos), as sets keyed by (query, clump, lane)."""
    nq = refout["nq"]
    inside = np.ones(w["nslots"], bool)
    inside[w["slot"][nq:]] = False                                   # a slot with a strand beyond the sample saw fewer visits on the CPU
    inside &= np.bincount(w["slot"][:nq], minlength=w["nslots"]) > 0
    minima_equal = bool(np.array_equal(best[inside], refout["best"][inside]))
    tq = runs["query0"][hits["task"] >> 4] + (hits["task"] & 15); tc = runs["clump"][hits["task"] >> 4]
    sl = w["slot"][tq]
    sel = inside[sl] & (sl % PARITY_SAMPLE_MOD == 0)
    g = np.zeros(int(sel.sum()), refout["hits"].dtype)
    g["query"] = tq[sel]; g["clump"] = tc[sel]
    for f in ("lane", "ed", "gap_q", "gap_r", "final_pos"):
        g[f] = hits[f][sel]
    g = g[np.lexsort((g["lane"], g["clump"], g["query"]))]
    rh = refout["hits"]; rh = rh[inside[w["slot"][rh["query"]]]]
    return {"minima_equal": minima_equal, "slots_compared": int(inside.sum()), "slots_with_hit": int((refout["best"][inside] != 0xFFFF).sum()),
            "hits_equal": bool(len(g) == len(rh) and np.array_equal(g, rh)), "hits_compared": int(len(rh)), "sample": "slot %% %d == 0" % PARITY_SAMPLE_MOD,
            "against": "reference kernels (oracle/_ref/libburstref.so) over the same bunches, %d of %d" % (refout["nb"], (len(w["qoff"]) - 1 + w["qbunch"] - 1) // w["qbunch"])}


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.config == "c3":
        wl = "configs[2] shape: %d x %d bp amplicon reads (0-%d substitutions, fwd+rc strands) per GPU vs %d MB synthetic mutation-tree DB (%d-column clumps of near-identical references), -i 0.97 budget %d, min selection, task list = reference bunch driver (QBUNCH 16 x ~60 candidate clumps per bunch)" % (
            args.reads, args.read_len, args.edits, args.db_mb, args.clump_len, args.budget)
    else:
        wl = ("north_star target: " if args.config == "target" else "configs[1]: ") + "%d x %d bp reads (exactly %d edits, LLsim model, fwd+rc strands) per GPU vs %d MB synthetic .edx-layout DB (%d-column clumps), -i 0.98 budget %d, BEST-style min selection, task list = reference bunch driver (QBUNCH 16 x bunch candidates)" % (
            args.reads, args.read_len, args.edits, args.db_mb, args.clump_len, args.edits)
    config = {"workload": wl,
        "reads_per_gpu": args.reads, "db_mb": args.db_mb, "sharding": "queries (DB replicated), no data-path collective", "numa_node_rank0": None,
        "l2": ("inputs (DB %d MB + tasks) exceed the 126 MB L2; no explicit flush" % args.db_mb) if args.db_mb > 126 else
              ("the %d MB DB fits the 126 MB L2 (as the reference's amplicon databases do); queries + runs + survivors of a step exceed it; no explicit flush" % args.db_mb)}

    if args.impl == "reference":
        if rank != 0:
            return
        w = build_workload(args, 0)
        cb, dt, _ = cpu_reference(args, w, steps=max(1, args.steps), warmup=min(args.warmup, 1))
        out = {"impl": "reference", "metric": "reads_per_sec", "value": cb["value"], "unit": "reads/s", "n_gpus": args.gpus,
               "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
               "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": config, "cpu_baseline": cb,
               "e2e": {"value": cb["value"], "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(out))
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the DP path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa_node(torch, local)             # host buffers (pinned) and the calling thread next to the GPU's PCIe root
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        # keep rank 0's stdout to the one JSON line: NCCL writes its version banner to fd 1 when the communicator is created
        sys.stdout.flush()
        saved = os.dup(1); os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush(); os.dup2(saved, 1); os.close(saved)
    from burst_b200.engine import Engine, MODE_MIN, RUN_DTYPE
    config["numa_node_rank0"] = numa

    w = build_workload(args, rank)
    stream = torch.cuda.Stream()
    eng = Engine(local, stream=stream.cuda_stream)
    eng.load_db(w["packed"], w["clump_len"])
    nq = len(w["qoff"]) - 1
    runs = np.ascontiguousarray(w["runs"], RUN_DTYPE)

    # pinned host staging for the e2e leg
    def pin(a):
        t = torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1)).pin_memory()
        return t.numpy().view(a.dtype).reshape(a.shape), t
    p_codes, k1 = pin(w["qcodes"]); p_off, k2 = pin(w["qoff"]); p_bud, k3 = pin(w["budget"]); p_slot, k4 = pin(w["slot"]); p_runs, k5 = pin(runs)
    h2d = p_codes.nbytes + p_off.nbytes + p_bud.nbytes + p_slot.nbytes + p_runs.nbytes

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- kernel-only: inputs resident in HBM ----------------
    eng.upload_runs(p_codes, p_off, p_bud, p_runs, slot=p_slot, nslots=w["nslots"])
    sampler = ClockSampler(local); sampler.start()           # nvidia-smi -lms 100 from the warm-up to the end of the e2e leg
    for _ in range(args.warmup):
        eng.run(MODE_MIN); eng.count()
    barrier()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    ms_filter = []
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(args.steps):
            eng.run(MODE_MIN)
        e1.record(stream)
    torch.cuda.synchronize()
    barrier()
    ms_total = e0.elapsed_time(e1)
    nhits = eng.count()
    st = eng.stats()
    hits, best = eng.download()
    found = int((best[:w["n_reads"]] <= args.budget).sum())
    # planted-read check: every read must be reported at the lane it was cut from
    tq = runs["query0"][hits["task"] >> 4] + (hits["task"] & 15); tc = runs["clump"][hits["task"] >> 4]
    rd = w["slot"][tq]
    ok = (tc == w["true_clump"][rd]) & (hits["lane"] == w["true_lane"][rd])
    planted = int(len(np.unique(rd[ok])))

    # ---------------- e2e: host buffers in, hits out, every step ----------------
    # the call a host driver makes: bg_align_runs_into() with its (pinned) query/run arrays and reusable (pinned) output
    # buffers; the library pipelines the host->device copies of the batch against its kernels and copies the hits back
    from burst_b200.engine import HIT_DTYPE
    p_hits, k6 = pin(np.zeros(max(nhits * 2, 1024), HIT_DTYPE)); p_best, k7 = pin(np.full(w["nslots"], 0xFFFF, np.uint16))
    p_pack, k8 = pin(Engine.pack4(w["qcodes"]))              # BG_Q_PACKED4 form of the same reads (two bases per byte, as the .edx stores references)

    def e2e_leg(codes_arg):
        for _ in range(min(args.warmup, 2)):
            p_best[:] = 0xFFFF
            eng.align_runs_into(codes_arg, p_off, p_bud, p_runs, p_hits, p_best, MODE_MIN, slot=p_slot, nslots=w["nslots"])
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            p_best[:] = 0xFFFF                               # per-slot minima carried in: none
            n = eng.align_runs_into(codes_arg, p_off, p_bud, p_runs, p_hits, p_best, MODE_MIN, slot=p_slot, nslots=w["nslots"])
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        assert n == nhits and np.array_equal(p_hits[:n], hits) and np.array_equal(p_best, best), "e2e path disagrees with the resident path"
        return dt, n * HIT_DTYPE.itemsize + p_best.nbytes

    e2e_bytes_s, d2h = e2e_leg(p_codes)                      # one code byte per base (burst.c's in-memory form)
    e2e_pack4_s, d2h = e2e_leg(("packed4", p_pack))          # nibble-packed strands through bg_align_runs_into()
    h2d_bytes_form = h2d + p_best.nbytes
    h2d_pack4 = h2d - p_codes.nbytes + p_pack.nbytes + p_best.nbytes

    # the compact form: every READ once at 2 bits per base, strands + runs derived on the device (bg_align_bunches_into) -- the headline e2e
    budget_r = np.zeros(w["n_reads"], np.uint16); budget_r[w["slot"]] = w["budget"]
    c_reads, k9 = pin(Engine.pack2(w["rcodes"])); c_len, k10 = pin(w["rlen"]); c_bud, k11 = pin(budget_r); c_strand, k12 = pin(w["strand"])
    c_coff, k13 = pin(w["cand_off"].astype(np.uint32)); c_cand, k14 = pin(w["cand"].astype(np.uint32))
    h2d = c_reads.nbytes + c_len.nbytes + c_bud.nbytes + c_strand.nbytes + c_coff.nbytes + c_cand.nbytes + p_best.nbytes

    def compact_leg():
        for _ in range(min(args.warmup, 2)):
            p_best[:] = 0xFFFF
            eng.align_bunches_into(c_reads, c_len, c_bud, c_strand, w["qbunch"], c_coff, c_cand, p_hits, p_best, MODE_MIN, packed2=True)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            p_best[:] = 0xFFFF
            n = eng.align_bunches_into(c_reads, c_len, c_bud, c_strand, w["qbunch"], c_coff, c_cand, p_hits, p_best, MODE_MIN, packed2=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        assert n == nhits and np.array_equal(p_hits[:n], hits) and np.array_equal(p_best, best), "compact e2e path disagrees with the resident path"
        return dt, n * HIT_DTYPE.itemsize + p_best.nbytes

    e2e_one_s, d2h = compact_leg()

    # Two batches in flight: a second context on the same GPU that borrows the database (bg_share_db), one host thread per context, the
    # same call.  Every step still copies its own inputs in and its hits + minima out inside the timed region; the copies of one step
    # travel behind the kernels of the other -- what the reference's thread team does over one shared database (burst.c:4050-4077) and
    # what the host driver does with its waves.  This is the headline e2e; the one-call-at-a-time figure is reported beside it.
    import threading
    NIF = max(1, args.inflight)
    lanes = [(eng, p_hits, p_best)]; keep = []
    for _ in range(NIF - 1):
        e_x = Engine(local); e_x.share_db(eng)
        h_x, ka = pin(np.zeros(len(p_hits), HIT_DTYPE)); b_x, kb = pin(np.full(w["nslots"], 0xFFFF, np.uint16))
        lanes.append((e_x, h_x, b_x)); keep += [ka, kb]

    def inflight_leg():
        got = [None] * NIF; err = []

        def work(k, count):
            e, hb, bb = lanes[k]
            try:
                for _ in range(count):
                    bb[:] = 0xFFFF
                    got[k] = e.align_bunches_into(c_reads, c_len, c_bud, c_strand, w["qbunch"], c_coff, c_cand, hb, bb, MODE_MIN, packed2=True)
            except Exception as ex:                          # noqa: BLE001
                err.append(ex)

        def both(counts):
            th = [threading.Thread(target=work, args=(k, counts[k])) for k in range(NIF) if counts[k]]
            for t in th:
                t.start()
            for t in th:
                t.join()
            if err:
                raise err[0]
        both([min(args.warmup, 2)] * NIF)
        barrier()
        t0 = time.perf_counter()
        both([(args.steps + NIF - 1 - k) // NIF for k in range(NIF)])
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        for k in range(NIF):
            if got[k] is not None:
                assert got[k] == nhits and np.array_equal(lanes[k][1][:nhits], hits) and np.array_equal(lanes[k][2], best), "in-flight e2e path disagrees with the resident path"
        return dt

    e2e_s = inflight_leg()
    for e_x, _, _ in lanes[1:]:
        e_x.close()
    clocks = sampler.stop()

    # ---------------- reference sharding (N > 1): the path's one collective, measured ----------------
    # SURVEY 8(e) / burst.c:4490-4519: the database is cut into N contiguous clump ranges, every GPU sees EVERY query; per step each GPU
    # runs filter + extend on its range, then ncclAllReduce(MIN) over the per-slot minima in place on the device, then the selection
    # against the combined minima.  Rank 0's read set (broadcast over NCCL) against the same database the replicated legs used, so the
    # union of the ranks' hits must equal rank 0's single-GPU result above -- checked here, at full size.
    ref_sharded = None
    if world > 1 and args.config in ("c2", "target"):
        from burst_b200 import sharded
        dev = torch.device("cuda", local)

        def bcast(a):
            a = np.ascontiguousarray(a)
            n = torch.tensor([a.nbytes if rank == 0 else 0], dtype=torch.int64, device=dev)
            dist.broadcast(n, 0)
            t = torch.from_numpy(a.view(np.uint8).reshape(-1)).to(dev) if rank == 0 else torch.empty(int(n.item()), dtype=torch.uint8, device=dev)
            dist.broadcast(t, 0)
            return t.cpu().numpy().view(a.dtype)
        # the replicated database must be the same on every rank (the generator seeds it independently of the rank): compare a checksum
        sample = w["packed"][:: max(1, len(w["packed"]) >> 20)].astype(np.int64)
        ck = torch.tensor([int(sample.sum()), -int(sample.sum()), len(w["clump_len"]), -len(w["clump_len"])], dtype=torch.int64, device=dev)
        dist.all_reduce(ck, op=dist.ReduceOp.MAX)
        same_db = int(ck[0]) == -int(ck[1]) and int(ck[2]) == -int(ck[3])
        if not same_db:
            ref_sharded = {"skipped": "ranks hold different databases"}
        else:
            g_codes = bcast(w["qcodes"]); g_off = bcast(w["qoff"]); g_bud = bcast(w["budget"]); g_slot = bcast(w["slot"]); g_runs = bcast(runs)
            eng2 = Engine(local, stream=stream.cuda_stream)
            with torch.cuda.stream(stream):
                drv = sharded.ReferenceSharded(eng2)
                lo, hi = drv.load_db(w["packed"], w["clump_len"])
                eng2.upload_runs(g_codes, g_off, g_bud, g_runs, slot=g_slot, nslots=w["nslots"])
                for _ in range(args.warmup):
                    drv.step_resident(MODE_MIN, w["nslots"])
                barrier()
                ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
                s0 = torch.cuda.Event(enable_timing=True); s1 = torch.cuda.Event(enable_timing=True)
                s0.record(stream)
                for k in range(args.steps):
                    gbest_t = drv.step_resident(MODE_MIN, w["nslots"], events=ev[k])
                s1.record(stream)
                torch.cuda.synchronize()
                barrier()
                rs_ms = s0.elapsed_time(s1) / args.steps
                ar_ms = sum(a.elapsed_time(b) for a, b in ev) / args.steps
                # the collective alone (ranks in step): the same tensor, back to back
                a0 = torch.cuda.Event(enable_timing=True); a1 = torch.cuda.Event(enable_timing=True)
                tmp = gbest_t.clone()
                dist.all_reduce(tmp, op=dist.ReduceOp.MIN); barrier()
                a0.record(stream)
                for _ in range(10):
                    dist.all_reduce(tmp, op=dist.ReduceOp.MIN)
                a1.record(stream)
                torch.cuda.synchronize()
                ar_alone = a0.elapsed_time(a1) / 10
                my_hits, _ = eng2.download()
                gbest = gbest_t.cpu().numpy().astype(np.uint16)
                allh = np.concatenate(sharded._all_gather_var(my_hits, None, dev))
            allh = allh[np.lexsort((allh["lane"], allh["task"]))]
            tt = torch.tensor([rs_ms, ar_ms, ar_alone], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            rs_ms, ar_ms, ar_alone = (float(x) for x in tt.cpu())
            ref_sharded = {"value": args.reads / (rs_ms / 1e3), "unit": "reads/s", "ms_per_step": rs_ms,
                           "allreduce_ms_in_step": ar_ms, "allreduce_ms_alone": ar_alone, "allreduce_bytes": int(w["nslots"]) * 4,
                           "clumps_this_rank": [int(lo), int(hi)], "db_mb_per_gpu": args.db_mb / world,
                           "workload": "rank 0's %d reads, every rank sees all of them; the %d MB database cut into %d contiguous clump ranges" % (args.reads, args.db_mb, world),
                           "protocol": "per step and rank: k_seedw + k_extend on the local clump range -> ncclAllReduce(MIN, u32 x reads) in place on the engine's stream -> k_select against the combined minima (burst.c:4490-4519); one 16-byte counter read-back (survivor-list overflow check) after the extend, no other host synchronisation",
                           "note": "allreduce_ms_in_step includes waiting for the slowest rank's extend; allreduce_ms_alone is the collective back to back with the ranks in step"}
            if rank == 0:
                hs = hits[np.lexsort((hits["lane"], hits["task"]))]
                ref_sharded["parity"] = {"hits_equal_single_gpu": bool(len(allh) == len(hs) and np.array_equal(allh, hs)), "minima_equal_single_gpu": bool(np.array_equal(gbest, best)),
                                         "hits": int(len(allh))}
            eng2.close()

    times = torch.tensor([ms_total, e2e_s * 1e3, e2e_bytes_s * 1e3, e2e_pack4_s * 1e3, e2e_one_s * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms, e2e_bytes_ms, e2e_pack4_ms, e2e_one_ms = (float(x) for x in times.cpu())
    ms_step = ms_total / args.steps
    total_reads = args.reads * world
    value = total_reads / (ms_step / 1e3)
    e2e_value = total_reads / (e2e_ms / 1e3 / args.steps)

    if rank == 0:
        peak, how = peaks()
        # Algorithmic bytes of one k_seedw launch: what the batched formulation must move at least once --
        # per clump visit of a bunch (run): the clump (8 B per column), its 16 B record and the 12 B run; per query: its 24 B
        # record and packed bases (len/2); 16 B per survivor written.  SURVEY 8(d)'s per-task figure (8*ClumpLen+len+16 per
        # (query, clump) pair) counts the clump once per query of the bunch, i.e. ~16x what one pass over it moves; it is
        # reported beside it as `achieved_per_task_8d`.
        run_cols = w["clump_len"][runs["clump"]].astype(np.int64)
        alg_bytes = float((8 * run_cols + 28).sum()) + nq * (24.0 + args.read_len / 2.0) + 16.0 * st["survivors"]
        alg_bytes_8d = float((8 * w["clump_len"][w["tasks"][:, 1]].astype(np.int64) + args.read_len + 16).sum()) + 12.0 * nhits
        filt_ms = st["ms_filter"]
        achieved = alg_bytes / (filt_ms / 1e3) / 1e9
        traffic = None
        try:    # dram__bytes_read.sum + dram__bytes_write.sum of one k_seedw launch of THIS workload, from the committed ncu pass
            tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            if tr["reads"] == args.reads and tr["db_mb"] == args.db_mb and tr["read_len"] == args.read_len:
                traffic = tr["k_seedw_dram_bytes_per_launch"]
        except Exception:
            pass
        out = {"metric": "reads_per_sec", "value": value, "unit": "reads/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32 (bit-parallel automata / packed DP keys; 8-bit reference semantics)",
               "data": "synthetic", "config": config,
               "e2e": {"value": e2e_value, "unit": "reads/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_ms / args.steps,
                       "call": "bg_align_bunches_into(): every read once at 2 bits per base + u16 length/budget, one u32 per strand, bunch -> candidate lists (u32), all in pinned host memory; the device derives both strands and the runs; hits + minima into pinned host buffers.  Several contexts per GPU (contexts_in_flight; bg_share_db: one database in HBM, kernel sequences gated one behind the other), one host thread each, steps dealt round-robin: every step's inputs are copied in and its results copied out inside the timed region, the copies of one step behind the kernels of the other",
                       "contexts_in_flight": NIF,
                       "one_call_at_a_time": {"value": total_reads / (e2e_one_ms / 1e3 / args.steps), "ms_per_step": e2e_one_ms / args.steps, "h2d_bytes_per_step": int(h2d),
                                              "call": "the same call from one host thread, one context: copy in, compute, copy out, then the next step (latency of one call)"},
                       "packed4_strands": {"value": total_reads / (e2e_pack4_ms / 1e3 / args.steps), "h2d_bytes_per_step": int(h2d_pack4), "ms_per_step": e2e_pack4_ms / args.steps,
                                           "call": "bg_align_runs_into(), both strands of every read nibble-packed (two bases per byte), offsets/budgets/slots per strand, 12-byte runs"},
                       "byte_codes": {"value": total_reads / (e2e_bytes_ms / 1e3 / args.steps), "h2d_bytes_per_step": int(h2d_bytes_form), "ms_per_step": e2e_bytes_ms / args.steps,
                                      "call": "the same call with one code byte per base"}},
               "gpu_launches": 15 * args.steps,
               "dp_gcups_nominal": st["nominal_cells"] * world / (ms_step / 1e3) / 1e9,
               "dp_gcups_executed": (st["filter_cells"] + st["band_cells"]) * world / (ms_step / 1e3) / 1e9,
               "seed_steps_per_s": st["seed_steps"] * world / (ms_step / 1e3),
               "work": {"tasks": st["tasks"], "survivors": st["survivors"], "hits": st["hits"], "nominal_cells": st["nominal_cells"],
                        "filter_cells": st["filter_cells"], "seed_steps": st["seed_steps"], "seed_queries": st["seed_queries"],
                        "seed_layout": "probe every %d columns, %d-base windows, %d-word filter per warp" % (st["seed_stride"], st["seed_window"], st["seed_words"]), "band_cells": st["band_cells"], "reads_found": found, "reads_at_planted_lane": planted,
                        "ms_filter": st["ms_filter"], "ms_extend": st["ms_extend"], "ms_select": st["ms_select"]},
               "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                            "kernel": "k_seedw" if st["seed_queries"] else "k_filter", "kernel_ms": filt_ms, "algorithmic_bytes_per_launch": alg_bytes,
                            "peak_source": "of " + how + " (MEASURED_PEAKS.json hbm_gbs, burst figure; fallback = 6650 GB/s of B200_PROFILING.md)",
                            "achieved_per_task_8d": alg_bytes_8d / (filt_ms / 1e3) / 1e9,
                            "note": "achieved = bytes one pass must move (clump once per run + queries once + survivors) / CUDA-event time of the k_seedw launch; the kernel streams each clump once for the <=16 queries of a bunch, so SURVEY 8(d)'s per-(query, clump) byte count (achieved_per_task_8d) exceeds what is physically read; the kernel is latency/issue bound, not HBM bound (profiles/)"},
               "integer_roofline": {"kernel": "k_extend (9 band classes, one binned survivor list)", "kernel_ms": st["ms_extend"],
                                    "dpx_thread_inst_per_s": 2.0 * st["band_cells"] / (st["ms_extend"] / 1e3),
                                    "dpx_peak_thread_inst_per_s": DPX_PEAK, "frac": 2.0 * st["band_cells"] / (st["ms_extend"] / 1e3) / DPX_PEAK,
                                    "note": "2 VIADDMNMX per band cell (select-with-tie-break of the packed pass-2 key); peak = VIADDMNMX.U32 issue rate measured on this pool's B200 by burst_b200/csrc/tools/pipe_microbench (profiles/r1d_pipe_microbench.txt: 571.6 G warp-inst/s at 1965 MHz, half the 4-per-clock issue rate: it shares the ALU pipe with the ~5 LOP3/SHF/VIMNMX each cell also needs)"},
               "clocks": clocks, "workload_gen_s": w["gen_s"]}
        if ref_sharded is not None:
            out["ref_sharded"] = ref_sharded
        if not args.no_cpu_baseline and world == 1:
            cb, _, refout = cpu_reference(args, w)
            out["cpu_baseline"] = cb
            if refout is not None:
                out["parity"] = parity_vs_reference(w, runs, hits, best, refout)
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    eng.close()


if __name__ == "__main__":
    main()
