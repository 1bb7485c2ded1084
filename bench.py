"""Benchmark of the TIP tri-graph encoder/decoder hot path (BASELINE.json metric:
"TIP-cat train step ms & typed-edge msgs/s at 1/2/4/8 B200; % HBM roofline").

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one rank per GPU under torchrun)
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU path (oracle port) on host cores

One "step" = one full-batch training step of TIP-cat (tip.py:24-30): encoder forward, typed negative
sampling, decoder + loss, backward, Adam -- on the synthetic polypharmacy-shape graph of SURVEY.md section 8(d)
(645 drugs, 19,081 proteins, 861 relations, ~8.28 M directed typed D-D edges).
value = typed-edge messages per second = 4 * E_directed / step_time (2 R-GCN layers + positive + negative
decoder pass, forward count; backward and Adam are in the time, not in the count).
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "typed_edge_msgs_per_s"
UNIT = "msgs/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mod", default="cat", choices=["cat", "add"])
    ap.add_argument("--shape", default="polypharmacy", choices=["polypharmacy", "small", "scaled"],
                    help="scaled = BASELINE.json config 4 (10k drugs, 100k proteins, 4k relations, ~50M directed D-D edges)")
    ap.add_argument("--model", default="tip", choices=["tip", "dd"],
                    help="dd = BASELINE.json config 3: the D-D-only R-GCN of test/dd_net_scalable.py, relation-sharded")
    ap.add_argument("--workload", default="train", choices=["train", "sweep"],
                    help="sweep = BASELINE.json config 5: the decoder-only inference sweep, all 645^2 drug pairs x 861 "
                         "relations (a step = one full prediction tensor through the C ABI)")
    ap.add_argument("--no-graph", action="store_true", help="do not capture the step in a CUDA graph")
    ap.add_argument("--cpu-sample-relations", type=int, default=160,
                    help="relations of the workload the CPU reference arm runs per step (its autograd backward costs "
                         "O(R*E*F): the full 861-relation step takes minutes, see profiles/*cpu_full_step*.json)")
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    return ap.parse_args()


def make_data(shape, mod="cat"):
    from tip_b200 import synth
    if shape == "small":
        return synth.make_tip_data(n_drug=200, n_prot=2000, n_rel=40, dd_undirected=60_000, pp_undirected=20_000,
                                   pd_edges=2_000, seed=1111), "synthetic small (debug) shape"
    if shape == "scaled":
        return synth.make_tip_data(**synth.SCALED, seed=1114), \
            "TIP-%s train step, scaled synthetic graph: 10000 drugs, 100000 proteins, 4000 relations" % mod
    return synth.make_tip_data(**synth.POLYPHARMACY, seed=1112), \
        "TIP-%s train step, synthetic polypharmacy shape: 645 drugs, 19081 proteins, 861 relations" % mod


def workload_config(workload, mod, data, n_gpus):
    """the `config` object: names the workload only, identical in both arms (b200 / reference)"""
    return {"workload": workload, "mod": mod, "directed_dd_edges": int(data["dd_train_idx"].shape[1]),
            "relations": int(data["n_dd_et"]), "drugs": int(data["n_drug"]), "proteins": int(data["n_prot"]),
            "parallelism": "relations sharded over %d GPU(s), P-P/P-D replicated" % n_gpus,
            "neg_sampler": "MT19937 bit-exact (numpy-compatible)",
            "l2": "per-step working set (index arrays + sampler stream, >0.6 GB) exceeds the 126 MB L2; no flush"}


def full_config_cpu_reference():
    """the one-off same-config measurement of the CPU reference arm (ONE full 861-relation step; tools/cpu_full_step.py),
    committed under profiles/ -- bounds the ratio taken on the sub-sampled arm"""
    best = None
    for name in ("r02a_cpu_full_step_box.json", "r02_cpu_full_step.json"):
        try:
            rec = json.load(open(os.path.join(ROOT, "profiles", name)))
            best = {"file": "profiles/" + name, "seconds_per_step": rec["seconds"], "value": rec["typed_edge_msgs_per_s"],
                    "cores": rec["cores"], "relations": rec["relations"], "directed_dd_edges": rec["directed_dd_edges"]}
            break
        except Exception:
            continue
    return best


def settings_for(mod):
    from tip_b200 import layers
    if mod == "cat":
        return layers.Setting(sp_rate=0.9, lr=0.01, prot_drug_dim=16, n_embed=48, n_hid1=32, n_hid2=16, num_base=32)
    return layers.Setting(sp_rate=0.9, lr=0.01, prot_drug_dim=64, n_embed=64, n_hid1=32, n_hid2=16, num_base=32)


# ----------------------------------------------------------------------------------------- clocks
class ClockSampler(object):
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 8 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 8 for n, v in zip(names, r[4:8]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ----------------------------------------------------------------------------------------- CPU reference arm
def cpu_reference_step_factory(data, mod, n_sample_rel):
    """The reference's CPU path for one training step, as the structural oracle (op-for-op restatement of
    src/layers.py + src/neg_sampling.py, incl. the per-relation Python loops), on the first `n_sample_rel`
    relations of the workload (a full step costs ~9 minutes on 8 cores)."""
    from oracle import neg_sampling_oracle as nso
    from oracle import tip_oracle as to
    from tip_b200 import layers
    n_rel = min(n_sample_rel, int(data["n_dd_et"]))
    rl = data["dd_train_range"][:n_rel].clone()
    e_s = int(rl[-1, 1])
    d = {"dd_train_idx": data["dd_train_idx"][:, :e_s].contiguous(), "dd_train_et": data["dd_train_et"][:e_s].contiguous(),
         "dd_train_range": rl, "pp_train_indices": data["pp_train_indices"], "dp_edge_index": data["dp_edge_index"],
         "d_norm": data["d_norm"]}
    torch.manual_seed(1111)
    s = settings_for(mod)
    enc = layers.FMEncoder("cpu", data["n_drug_feat"], n_rel, data["n_prot"], data["n_prot"], data["n_drug"],
                           s.prot_drug_dim, s.num_base, s.n_embed, s.n_hid1, s.n_hid2, mod=mod)
    dec = layers.MultiInnerProductDecoder(s.n_hid2, n_rel)
    params = {"encoder." + n: p.detach().clone().requires_grad_(True) for n, p in enc.named_parameters()}
    params["decoder.weight"] = dec.weight.detach().clone().requires_grad_(True)
    orc = to.TipOracle(params, data["n_drug"], data["n_prot"], mod=mod, structural=True)
    opt = torch.optim.Adam(list(params.values()), lr=s.lr)
    mt = nso.MT19937(1111)
    pos_np, rl_np = d["dd_train_idx"].numpy(), rl.numpy()

    def step(phases=None):
        t = [time.perf_counter()]
        opt.zero_grad()
        neg = torch.from_numpy(nso.typed_negative_sampling(mt, pos_np, data["n_drug"], rl_np))
        t.append(time.perf_counter())
        loss, _ = orc.loss(d, neg)
        t.append(time.perf_counter())
        loss.backward()
        t.append(time.perf_counter())
        opt.step()
        t.append(time.perf_counter())
        if phases is not None:
            for name, a, b in zip(("neg_sampling", "forward", "backward", "adam"), t[:-1], t[1:]):
                phases[name] = b - a
        return float(loss.detach())

    return step, e_s, n_rel


def time_cpu_reference(data, mod, n_sample_rel, steps, warmup):
    step, e_s, n_rel = cpu_reference_step_factory(data, mod, n_sample_rel)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return {"value": 4.0 * e_s / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"first {n_rel} of {int(data['n_dd_et'])} relations ({e_s} directed D-D edges) + full P-P/P-D graphs; "
                      f"{steps} step(s) of structural oracle fwd+neg-sampling+bwd+Adam, {dt:.2f} s/step",
            "ms_per_step": dt * 1e3}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    data, workload = make_data(args.shape, args.mod)
    res = time_cpu_reference(data, args.mod, args.cpu_sample_relations, args.steps, args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(workload, args.mod, data, args.gpus),
            "cpu_baseline": dict({k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
                                 full_config=full_config_cpu_reference()),
            "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------- B200 arm
def count_library_launches(step_fn):
    """kernels of libtipb200 (namespace tipb::) launched by one eager step, counted with CUPTI."""
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        step_fn()
        torch.cuda.synchronize()
    mine = total = 0
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA and not ev.name.startswith("Memcpy") and not ev.name.startswith("Memset"):
            total += 1
            if "tipb" in ev.name:
                mine += 1
    return mine, total


def profile_serial_steps(model, step_fn, n_steps=3):
    """Per-kernel device time of a training step with EVERY kernel on one stream (no side streams, no prefetch), so no
    two kernels overlap and a kernel's duration is its own: CUPTI (torch.profiler) over `n_steps` eager steps after one
    untimed step in that mode.  -> list of steps, each a list of (kernel name, duration us) in launch order."""
    from torch.profiler import ProfilerActivity, profile
    from tip_b200 import layers, neg_sampling as ns
    layers.SERIAL_STREAMS = True
    ns.set_prefetch(False)
    try:
        step_fn()
        torch.cuda.synchronize()
        steps = []
        for _ in range(n_steps):
            with profile(activities=[ProfilerActivity.CUDA]) as prof:
                step_fn()
                torch.cuda.synchronize()
            evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA
                   and not e.name.startswith("Memcpy") and not e.name.startswith("Memset")]
            evs.sort(key=lambda e: e.time_range.start)
            steps.append([(e.name, float(e.time_range.end - e.time_range.start)) for e in evs])
    finally:
        layers.SERIAL_STREAMS = False
        ns.set_prefetch(True)
    return steps


def roofline_block(model, serial_steps, step_s, peaks):
    """`roofline` of the bench line: the kernel with the largest share of the (serialised) step, its (A)/t against the
    measured HBM peak next to its ncu DRAM traffic, the same per kernel family, and the step-level figure"""
    from tools import roofline as rf
    W = rf.workload_dims(model)
    peak = float(peaks.get("hbm_gbs", 6650.0))
    per_kernel, per_family = {}, {}
    for kernels in serial_steps:
        sm = rf.StepModel(W)
        for name, us in kernels:
            fam, alg, formula = sm.label(name)
            short = rf.short_name(name) if "tipb" in name else "(torch) " + rf.short_name(name)[:48]
            k = per_kernel.setdefault(short, {"us": 0.0, "launches": 0, "alg": 0, "formula": formula, "family": fam})
            k["us"] += us
            k["launches"] += 1
            k["alg"] += alg or 0
            f = per_family.setdefault(fam, {"us": 0.0, "alg": 0})
            f["us"] += us
            f["alg"] += alg or 0
    n = float(len(serial_steps))
    total_us = sum(k["us"] for k in per_kernel.values()) / n
    mine = {k: v for k, v in per_kernel.items() if not k.startswith("(torch)")}
    if not mine:    # CUPTI is taken (the run is itself under ncu / another profiler): no per-kernel times, no roofline claim
        return {"bound": "hbm", "kernel": None, "achieved": None, "peak": peak, "unit": "GB/s", "frac": None, "traffic": None,
                "note": "no CUPTI kernel records (profiler attached): a line printed under a profiler is not a bench value"}
    dom = max(mine, key=lambda k: mine[k]["us"])
    dk = mine[dom]
    launches = dk["launches"] / n
    k_us = dk["us"] / dk["launches"]                       # average launch duration
    alg = dk["alg"] / dk["launches"] if dk["alg"] else None
    traffic, traffic_src = None, None
    try:    # dram__bytes_read.sum + dram__bytes_write.sum per launch of this kernel, from the committed ncu capture
        t = json.load(open(os.path.join(ROOT, "profiles", "kernel_traffic.json")))
        if dom in t["kernels"]:
            traffic, traffic_src = t["kernels"][dom]["traffic_bytes"], t["source"]
    except Exception:
        pass
    achieved = alg / (k_us * 1e-6) / 1e9 if alg else None
    fam_rows = []
    for fam, v in sorted(per_family.items(), key=lambda kv: -kv[1]["us"]):
        us = v["us"] / n
        a = v["alg"] / n
        fam_rows.append({"family": fam, "us": round(us, 1), "share": round(us / total_us, 4),
                         "algorithmic_bytes": int(a) if a else None,
                         "frac": round(a / (us * 1e-6) / 1e9 / peak, 3) if a else None})
    kern_rows = [{"kernel": k, "us_per_step": round(v["us"] / n, 1), "launches_per_step": v["launches"] / n,
                  "frac": round(v["alg"] / v["us"] * 1e6 / 1e9 / peak, 3) if v["alg"] else None}
                 for k, v in sorted(per_kernel.items(), key=lambda kv: -kv[1]["us"])[:16]]
    step_alg = rf.step_algorithmic_bytes(W)
    block = {"bound": "hbm", "kernel": dom, "kernel_family": dk["family"], "kernel_share_of_step": round(dk["us"] / n / total_us, 4),
             "kernel_us": round(k_us, 2), "kernel_launches_per_step": launches,
             "achieved": achieved, "peak": peak, "peak_source": "MEASURED_PEAKS.json (burst copy)" if peaks else "fallback 6.65 TB/s",
             "unit": "GB/s", "frac": achieved / peak if achieved else None, "algorithmic_bytes": int(alg) if alg else None,
             "formula": dk["formula"], "traffic": traffic, "traffic_source": traffic_src,
             "dram_frac": traffic / (k_us * 1e-6) / 1e9 / peak if traffic else None,
             "how": "kernel durations: CUPTI over %d eager steps with every kernel on ONE stream (no overlap); the kernel "
                    "is the library kernel with the largest summed time in that step" % len(serial_steps),
             "serial_step_us": round(total_us, 1),
             "step": {"algorithmic_bytes": int(step_alg), "frac_of_timed_step": step_alg / step_s / 1e9 / peak,
                      "what": "SURVEY 8(d): sum of (A) over the step / the timed (overlapped, CUDA-graph) step time"},
             "families": fam_rows, "kernels": kern_rows,
             "note": "(A) counts gathered payload rows, which this path serves from shared memory; where frac > 1 the "
                     "kernel is not HBM-bound (issue / shared-memory bound) and dram_frac is the DRAM-pin figure"}
    return block


def run_b200_arm(args):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs CUDA devices (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import datetime
        opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=120), pg_options=opts)
    from tip_b200 import layers, neg_sampling as ns

    data, workload = make_data(args.shape, args.mod)
    e_total = int(data["dd_train_idx"].shape[1])
    torch.manual_seed(1111)
    ns.seed(1111, dev)
    if args.model == "dd":
        workload = workload.replace("TIP-%s train step" % args.mod, "D-D-only R-GCN (test/dd_net_scalable.py) train step")
    if world > 1 or args.model == "dd":
        from tip_b200 import parallel
        # the sampler's all-gather runs on a side stream: give it its own communicator so that it does not queue
        # behind the encoder's all-reduces on NCCL's stream
        coll = parallel._Collective(world, sampler_group=dist.new_group() if world > 1 else None)
        cls = parallel.ShardedDDNet if args.model == "dd" else parallel.ShardedTIP
        model = cls(settings_for(args.mod), dev, mod=args.mod, data=data, rank=rank, world=world, collective=coll,
                    defer_loss_reduce=world > 1)
    else:
        model = layers.TIP(settings_for(args.mod), dev, mod=args.mod, data=data)
    sharded = hasattr(model, "refresh_shard")
    if os.environ.get("TIPB_BENCH_TORCH_ADAM") == "1":
        opt = torch.optim.Adam(model.parameters(), lr=model.settings.lr, capturable=True, fused=True)
    else:       # tip.py:21 `torch.optim.Adam(model.parameters(), lr)` as one launch of the library (csrc/adam.cu)
        from tip_b200 import optim
        opt = optim.Adam(model.parameters(), lr=model.settings.lr)

    one = torch.ones((), dtype=torch.float32, device=dev)     # d(loss)/d(loss): autograd would fill a fresh one per step

    def step():
        opt.zero_grad(set_to_none=True)
        loss = model(check_status=False)
        loss.backward(one)
        opt.step()
        ns.join_prefetch(dev)        # the next step's MT19937 words were generated on a side stream meanwhile
        return model.last_loss if getattr(model, "defer_loss_reduce", False) else loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up on a side stream (fills every cache / workspace; the recipe CUDA-graph capture needs)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(max(args.warmup, 3)):
            loss_value = float(step().detach())
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    assert int(ns.last_status(dev)) == 0, "negative sampler ran out of pre-generated words"

    graph = None
    main_prio = int(os.environ.get("TIPB_BENCH_MAIN_PRIORITY", "-1"))
    static_loss = None
    if not args.no_graph:
        try:
            model.embeddings = None          # drop the last eager autograd graph before capturing
            opt.zero_grad(set_to_none=True)
            graph = torch.cuda.CUDAGraph()
            # the main chain (P-P encoder, R-GCN, pair pass, backward, collectives) is captured on a high-priority stream, the
            # sampler's side stream has the default priority (tip_b200/layers.py: SIDE_PRIORITY)
            cap = torch.cuda.Stream(device=dev, priority=main_prio) if main_prio != 0 else None
            with torch.cuda.graph(graph, stream=cap):
                static_loss = step().detach()
            for _ in range(2):
                graph.replay()
            torch.cuda.synchronize()
        except Exception as exc:  # capture is an optimisation; eager is the fallback (still the CUDA path)
            if rank == 0:
                print(f"[bench] CUDA graph capture failed, running eagerly: {exc}", file=sys.stderr)
            graph = None
            torch.cuda.synchronize()

    run_step = (lambda: graph.replay()) if graph is not None else step

    # ---- timed region: K steps, device events, max over ranks
    clocks = ClockSampler(local)
    clocks.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        run_step()
    e1.record()
    barrier()
    clock_info = clocks.stop()
    elapsed = e0.elapsed_time(e1) * 1e-3
    if world > 1:
        t = torch.tensor([elapsed], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed = float(t)
    assert int(ns.last_status(dev)) == 0
    step_s = elapsed / args.steps
    value = 4.0 * e_total / step_s

    # ---- end-to-end: every step copies the step's inputs (the graph tensors, int64 as the reference API has them)
    #      from pinned host memory, rebuilds every index structure from them, trains one step and reads the loss back
    if static_loss is not None:
        loss_value = float(static_loss)
    try:
        e2e = measure_e2e(model, opt, data, args.steps, e_total, world, sharded)
    except Exception as exc:      # keep the device-timed line if the end-to-end loop fails (the same way on every rank)
        if world == 1:
            raise
        print(f"[bench] rank {rank}: end-to-end measurement failed: {exc!r}", file=sys.stderr)
        e2e = None
    launches, launches_all = count_library_launches(step)   # every rank runs it: the step contains collectives

    serial = profile_serial_steps(model, step, 3)            # every rank runs it: the step contains collectives
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        roofline = roofline_block(model, serial, step_s, peaks)
        cpu = None
        if not args.skip_cpu_baseline and world == 1:
            torch.set_num_threads(os.cpu_count() or 1)
            cpu = time_cpu_reference(data, args.mod, args.cpu_sample_relations, 1, 1)
            cpu = dict({k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")},
                       full_config=full_config_cpu_reference())
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": step_s * 1e3, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": workload_config(workload, args.mod, data, world), "cuda_graph": graph is not None,
                "stream_priority": {"step": main_prio if graph is not None else 0, "sampler_side_stream": layers.SIDE_PRIORITY},
                "clocks": clock_info, "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
                "gpu_launches": launches * args.steps, "gpu_launches_per_step": launches,
                "all_cuda_kernels_per_step": launches_all, "loss": loss_value}
        print(json.dumps(line), flush=True)
    if world > 1:
        # Tearing NCCL down while captured CUDA graphs still reference the communicator can block; every
        # rank is past its last collective here, so synchronise on the device and leave without the teardown.
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def measure_e2e(model, opt, data, steps, e_total, world=1, sharded=False):
    """every rank copies ITS inputs of the step from pinned host memory -- the graph tensors the path reads, int64 as the
    reference API has them (a sharded rank: its relations' edges only; `dd_train_et` is not copied: like the reference's
    MyRGCNConv2 the path takes relation membership from range_list) -- rebuilds its index structures from them, trains
    one step and reads the loss back; wall clock between barriers, max over ranks"""
    import torch.distributed as dist
    dev = model.device
    d = model.data
    one = torch.ones((), dtype=torch.float32, device=dev)
    pairs = []          # (device tensor, pinned host tensor)
    if sharded:
        pairs.append((model.local_idx, data["dd_train_idx"][:, model.e_lo:model.e_hi].contiguous().pin_memory()))
    else:
        pairs.append((d.dd_train_idx, data["dd_train_idx"].contiguous().pin_memory()))
    for k in ("dd_train_range", "pp_train_indices", "dp_edge_index", "d_norm"):
        pairs.append((getattr(d, k), data[k].to(getattr(d, k).dtype).contiguous().pin_memory()))
    h2d = sum(h.numel() * h.element_size() for _, h in pairs)

    def e2e_step():
        for dst, src in pairs:                # host -> device, in place: the version bump makes every cached typed CSR /
            dst.copy_(src, non_blocking=True)         # bitmap rebuild itself (into the same buffers)
        if sharded:
            model.refresh_shard()
        opt.zero_grad(set_to_none=True)
        loss = model(check_status=False)
        loss.backward(one)
        opt.step()
        out = model.last_loss if getattr(model, "defer_loss_reduce", False) else loss
        return out.item()                     # device -> host

    def fence():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(2):
        e2e_step()
    fence()
    t0 = time.perf_counter()
    for _ in range(steps):
        e2e_step()
    fence()
    dt = (time.perf_counter() - t0) / steps
    if world > 1:
        t = torch.tensor([dt], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        h = torch.tensor([h2d], device=dev, dtype=torch.float64)
        dist.all_reduce(h, op=dist.ReduceOp.SUM)
        dt, h2d = float(t), int(h)
    return {"value": 4.0 * e_total / dt, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4 * world,
            "ms_per_step": dt * 1e3,
            "what": "per rank: its graph tensors pinned-host -> device (int64; a sharded rank copies its relations' edges "
                    "only), all its typed CSRs / bitmaps / pair tables rebuilt, one train step, loss.item(); wall clock "
                    "between barriers, max over ranks; bytes summed over ranks"}


def run_sweep(args):
    """BASELINE.json config 5: the full prediction tensor [861, 645, 645] (1.43 GB of fp32 scores) per step, through
    tipb_decoder_sweep on a preallocated output; L2 flushed between steps; roofline = the DRAM write stream."""
    from tip_b200 import _lib
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    if int(os.environ.get("RANK", "0")) != 0:
        return                                   # one relation-independent kernel: replicas only, rank 0 reports
    torch.manual_seed(0)
    n, r, dim = 645, 861, 16
    z = torch.randn(n, dim, device=dev)
    w = torch.randn(r, dim, device=dev) * 0.25
    z_host, w_host = z.cpu().pin_memory(), w.cpu().pin_memory()
    out = torch.empty((r, n, n), dtype=torch.float32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    L = _lib.lib()

    def step():
        _lib.check(L.tipb_decoder_sweep(z.data_ptr(), w.data_ptr(), n, r, dim, 1, out.data_ptr(), _lib.stream()), "decoder_sweep")

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    clocks = ClockSampler(dev.index)
    clocks.start()
    total = 0.0
    for _ in range(args.steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step()
        e1.record()
        e1.synchronize()
        total += e0.elapsed_time(e1) * 1e-3
    clock_info = clocks.stop()
    assert L.tipb_decoder_sweep_status() == 0
    t = total / args.steps
    # end to end: embeddings and relation weights from pinned host memory, one score row block read back
    probe = torch.empty(n, dtype=torch.float32).pin_memory()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        z.copy_(z_host, non_blocking=True)
        w.copy_(w_host, non_blocking=True)
        step()
        probe.copy_(out[r - 1, n - 1], non_blocking=True)
        torch.cuda.synchronize()
    te = (time.perf_counter() - t0) / args.steps
    rel = torch.tensor([0, 17, 430, 860], device=dev)
    ref = torch.sigmoid(torch.einsum("ik,rk,jk->rij", z.double(), w[rel].double(), z.double()))
    err = float((out[rel].double() - ref).abs().max())
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    nbytes = out.numel() * 4
    line = {"metric": "decoder_sweep_scores_per_s", "value": out.numel() / t, "unit": "scores/s", "n_gpus": 1,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": t * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32 (three bf16 pieces per operand on tcgen05, fp32 accumulation)",
            "data": "synthetic", "config": {"workload": "decoder-only inference sweep: all 645^2 drug pairs x 861 relations "
                                            "(BASELINE.json config 5)", "scores": out.numel(), "l2": "256 MB flush between steps"},
            "clocks": clock_info,
            "roofline": {"bound": "hbm", "kernel": "k_decoder_sweep_tc<16>", "achieved": nbytes / t / 1e9, "peak": peak,
                         "peak_source": "MEASURED_PEAKS.json (burst copy)" if peaks else "fallback 6.65 TB/s", "unit": "GB/s",
                         "frac": nbytes / t / 1e9 / peak, "algorithmic_bytes": nbytes, "formula": "4 B per score written",
                         "traffic": None},
            "cpu_baseline": None,
            "e2e": {"value": out.numel() / te, "unit": "scores/s", "h2d_bytes_per_step": (z.numel() + w.numel()) * 4,
                    "d2h_bytes_per_step": n * 4, "ms_per_step": te * 1e3},
            "gpu_launches": args.steps, "gpu_launches_per_step": 1, "max_abs_err_vs_float64": err}
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    if args.workload == "sweep" and args.impl != "reference":
        run_sweep(args)
        return
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
