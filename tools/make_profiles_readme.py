"""Regenerates profiles/README.md (round 2) from the committed files under profiles/r02 (and r01 for the comparison columns).
Every number in the README is read from a file named next to it; a file that is missing yields "n/a" instead of a guess."""
import csv, glob, json, os, re
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R1 = os.path.join(ROOT, "profiles", "r01")
R2 = os.path.join(ROOT, "profiles", "r02")


def load(path):
    try:
        txt = open(path).read()
        lines = [l for l in txt.splitlines() if l.startswith("{")]
        return json.loads(lines[-1]) if lines else None
    except Exception:
        return None


def first(*names):
    for n in names:
        d = load(os.path.join(R2, n))
        if d:
            return d, "r02/" + n
    return None, None


def f(x, fmt="{:.1f}"):
    try:
        return fmt.format(x)
    except Exception:
        return "n/a"


def launch_shares(path, top=8):
    try:
        rows = list(csv.reader(open(path)))
        hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
        ix = {h: i for i, h in enumerate(rows[hi])}
        agg, n = {}, {}
        for r in rows[hi + 1:]:
            if len(r) < len(rows[hi]):
                continue
            name = re.sub(r"<.*", "", r[ix["Kernel Name"]].replace("void ", "").replace("ksg::", "")).split("(")[0].strip()
            name = "cub::DeviceRadixSort*" if "RadixSort" in name else ("cub::DeviceScan/Select*" if ("DeviceScan" in name or "DeviceSelect" in name or "DeviceCompact" in name) else name)
            v = float(r[ix["Metric Value"]].replace(",", ""))
            agg[name] = agg.get(name, 0.0) + v
            n[name] = n.get(name, 0) + 1
        tot = sum(agg.values())
        return [(k, 100.0 * v / tot, n[k]) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:top]], tot
    except Exception:
        return [], 0.0


def ncu_rows(path):
    """-> list of dicts (one per captured launch) of the raw page"""
    try:
        rows = [r for r in csv.reader(open(path)) if r]
        hdr, units = rows[0], rows[1]
        sc = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1e-3, "ms": 1.0, "ns": 1e-6, "s": 1e3}
        out = []
        for vals in rows[2:]:
            d = {}
            for h, u, v in zip(hdr, units, vals):
                try:
                    d[h] = float(v.replace(",", "")) * sc.get(u, 1)
                except ValueError:
                    d[h] = v
            out.append(d)
        return out
    except Exception:
        return []


fin, fin_src = first("bench_final_21.json", "bench_final_17.json", "bench_final_13.json", "bench_full_12.json", "bench_full_10.json")
ref, ref_src = first("bench_fast5_reference_21.json", "bench_fast5_reference_17.json", "bench_fast5_reference_13.json", "bench_fast5_reference_10.json")
r1 = load(os.path.join(R1, "bench_fast5.json"))
r1m = load(os.path.join(R1, "bench_merged2.json"))
m2 = (fin or {}).get("workloads", {}).get("merged2")
ph = fin["roofline"]["phase_ms_per_frame"]
tl = fin["roofline"].get("solve_kernel_timeline_last_profiled_frame") or {}
cpu = fin.get("cpu_baseline") or (load(os.path.join(R2, "bench_full_10.json")) or {}).get("cpu_baseline") or {}
cpu_m = (m2 or {}).get("cpu_baseline") or ((load(os.path.join(R2, "bench_full_10.json")) or {}).get("workloads", {}).get("merged2", {}) or {}).get("cpu_baseline") or {}
shim = fin.get("e2e_shim") or (load(os.path.join(R2, "bench_full_10.json")) or {}).get("e2e_shim") or {}
shim_m = (m2 or {}).get("e2e_shim") or ((load(os.path.join(R2, "bench_full_10.json")) or {}).get("workloads", {}).get("merged2", {}) or {}).get("e2e_shim") or {}
ms = fin.get("multi_sequence") or {}
out = []
w = out.append
w("# profiles/ — measured evidence (B200, sm_100a, CUDA 12.9, driver 580)\n")
w("Everything here was produced on the `gpurun` B200 boxes; a number taken under `ncu` is never a bench value.  `r02/` = this round,")
w("`r01/` = round 1 (its own `README.md` inside).  This file is generated from the committed files by `tools/make_profiles_readme.py`;")
w("every number names the file it comes from.\n")
w(f"## Headline (`python bench.py`: `fast5` = 640x480 depth+label stream, 5 cm voxels, 21 classes, `fast`; {fin['steps']} steps after {fin['warmup']} warm-up) — `{fin_src}`\n")
w("| arm | frames/s | ms/frame | round 1 |")
w("|---|---|---|---|")
w(f"| ours, frames resident in HBM (`value`, CUDA events on the launching stream) | **{f(fin['value'], '{:.0f}')}** | {f(fin['ms_per_step'], '{:.3f}')} | {f(r1['value'], '{:.0f}') if r1 else 'n/a'} |")
w(f"| ours, end to end from page-locked host frames, pipelined (`e2e.value`: `ksg_integrate_depth_async` + `ksg_wait_frame`, H2D of 1.5 MB/frame + counter read-back per step, wall clock) | **{f(fin['e2e']['value'], '{:.0f}')}** | {f(1e3 / fin['e2e']['value'], '{:.3f}')} | {f(r1['e2e']['value'], '{:.0f}') if r1 else 'n/a'} |")
w(f"| ours, end to end, synchronous call per frame (`e2e.sync_value`: `ksg_integrate_depth`, the reference's calling convention) | {f(fin['e2e']['sync_value'], '{:.0f}')} | {f(1e3 / fin['e2e']['sync_value'], '{:.3f}')} | — |")
if ref:
    w(f"| reference arm `bench.py --impl reference` (CPU: {ref['cpu_baseline'].get('kind')} @ {ref['cpu_baseline'].get('cores')} threads, fastest of the calibrated variants) — `{ref_src}` | {f(ref['value'])} | {f(ref['ms_per_step'])} | 36.3 |")
if cpu:
    w(f"| `cpu_baseline` inside our run ({cpu.get('kind')} @ {cpu.get('cores')} threads of {cpu.get('host_cores')}) | {f(cpu.get('value'))} | {f(1e3 / cpu['value']) if cpu.get('value') else 'n/a'} | 40.1 |")
if ms:
    w(f"| {ms.get('sequences')} independent sequences on ONE GPU (`multi_sequence`: how far one stream is from filling the machine) | {f(ms.get('value'), '{:.0f}')} | — | — |")
ee, le = (shim or {}).get("eager") or {}, (shim or {}).get("lazy") or {}
w(f"| the reference's own call through the C++ drop-in classes (`e2e_shim`: `SemanticTsdfIntegratorFactory::create` + `integratePointCloud` on host clouds), host layers refreshed every call (eager, the reference's contract) | {f(ee.get('fps'))} | {f(1e3 / ee['fps'], '{:.2f}') if ee.get('fps') else 'n/a'} | not measured |")
w(f"| same, host layers refreshed once at the end (lazy; a full export of the map, inside the span) | {f(le.get('fps'))} | — | — |")
ck = fin.get("clocks") or {}
w(f"\nSM clock {ck.get('sm_mhz')} MHz (max {ck.get('sm_max_mhz')}), throttle reasons {ck.get('reasons')}; {fin.get('gpu_launches', 0) / fin['steps']:.0f} own kernel launches and "
  f"{fin.get('library_calls', 0) / fin['steps']:.0f} library calls per frame (round 1: 17 + 3 CUB calls = 35 launches, 3-4 blocking read-backs).")
if ref:
    w(f"End-to-end speed-up over the reference CPU arm measured on the same box: {fin['e2e']['value'] / ref['value']:.0f}x.\n")
if m2:
    ee, le = (shim_m or {}).get("eager") or {}, (shim_m or {}).get("lazy") or {}
    w(f"`merged2` (BASELINE configs[2]: 640x480, 2 cm, 21 classes, `merged`, the REFERENCE's `unordered_map` bundle order = the default; ≈31 M voxel updates per frame), "
      f"`workloads.merged2` of the same line: **{f(m2['value'])} frames/s** = {f(m2['mvoxel_updates_per_s'], '{:.0f}')} Mvoxel-updates/s resident, "
      f"{f(m2['e2e']['value'])} end to end ({f(m2['e2e']['sync_value'])} synchronous); round 1: {f(r1m['value']) if r1m else 'n/a'} in a non-reference order.  "
      f"CPU port at its best thread count ({cpu_m.get('cores')}): {f(cpu_m.get('value'), '{:.2f}')} frames/s.  Through the C++ classes: eager {f(ee.get('fps'), '{:.1f}')}, lazy {f(le.get('fps'), '{:.1f}')} frames/s "
      f"(eager `merged` still copies whole updated blocks; the update log is `fast`-only).\n")

w("## Where a `fast5` frame goes\n")
w(f"CUDA events inside the library + `clock64` marks inside the persistent kernel (`roofline.phase_ms_per_frame` and `solve_kernel_timeline_last_profiled_frame` of `{fin_src}`; the profiling pass is separate from the timed pass and slower: its in-kernel probes cost ≈0.1 ms).\n")
r1p = (r1 or {}).get("roofline", {}).get("phase_ms_per_frame", {})
w("| phase | ms (r02) | ms (r01) | what runs (r02) |")
w("|---|---|---|---|")
desc = {"classify+start_set": "3 memsets, `k_fast_count`, `k_fast_classify`, `k_fast_start_eval3`, then inside `k_fast_solve3`: shared-slot sort, compaction, ray set-up",
        "fixpoint|bundling": "`k_fast_solve3`: observed-set sweeps (grid barrier between sweeps, no host read-back)",
        "ray_emit": "`k_fast_solve3`: table commit, block allocation + init", "record_sort": "`k_fast_solve3`: records scattered to their tiles (no sort)",
        "alloc+tile_heads": "(folded into the solve kernel)", "tile_apply": "`k_tile_apply_fast<TMA,1>` (shared-memory counting sort per tile, TMA bulk load/store)", "frame": "5 launches"}
for k in ("classify+start_set", "fixpoint|bundling", "ray_emit", "record_sort", "alloc+tile_heads", "tile_apply", "frame"):
    w(f"| {k} | {f(ph.get(k), '{:.3f}')} | {f(r1p.get(k), '{:.3f}')} | {desc[k]} |")
if tl:
    w(f"\nInside `k_fast_solve3` (last profiled frame, µs): shared-slot filing {f((tl.get('phase0_us') or {}).get('file_shared_slot_visitors'))} + sort {f((tl.get('phase0_us') or {}).get('sort_shared_slots'))} + "
      f"scan {f((tl.get('phase0_us') or {}).get('scan_cast_counts'))} + compaction {f((tl.get('phase0_us') or {}).get('compaction'))}; ray set-up {f(tl.get('ray_setup_us'))}; "
      f"{tl.get('sweeps')} sweeps {[round(x) for x in tl.get('sweep_us', [])]}; commit {f(tl.get('commit_emit_us'))}; tile alloc + block init {f(tl.get('tile_alloc_block_init_us'))}; scatter {f(tl.get('scatter_us'))}; "
      f"kernel {f(tl.get('solve_kernel_us'))}.  {(tl.get('debug') or {}).get('rays')} rays, {(tl.get('debug') or {}).get('ray_evals')} ray evaluations, {(tl.get('debug') or {}).get('ray_evals_that_changed')} of them changed something.")
sh, tot = launch_shares(os.path.join(R2, "launches_fast5_17.csv"))
src_l = "r02/launches_fast5_17.csv"
if not sh:
    sh, tot = launch_shares(os.path.join(R2, "launches_fast5_13.csv")); src_l = "r02/launches_fast5_13.csv"
if sh:
    w(f"\nncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`, cold caches, serialised - shares only) `{src_l}`: " +
      ", ".join(f"`{k}` {p:.0f} %" for k, p, _ in sh[:6] if k) + ".  The solve kernel's share agrees with the event-timed phases (everything but the pre-kernels and the tile apply).")
nr = ncu_rows(os.path.join(R2, "prof_solve3_fast5.raw.csv"))
if nr:
    d = nr[0]
    w(f"\n`ncu --set full` of `k_fast_solve3` (`r02/prof_solve3_fast5.details.txt`, `.raw.csv`; rank-group default of that commit): duration {f(d.get('gpu__time_duration.sum'), '{:.3f}')} ms, "
      f"DRAM {f((d.get('dram__bytes_read.sum', 0) + d.get('dram__bytes_write.sum', 0)) / 1e6)} MB per launch (algorithmic bytes of the frame: {f(fin['roofline']['algorithmic_bytes_per_launch'] / 1e6)} MB), "
      "L2 throughput 3 %, executed IPC 0.25, 93 % of the scheduler cycles without an eligible warp, and 80 % of all warp stall samples are the CTA barrier "
      "in front of the grid barrier: the kernel is a chain of ≈20 grid-wide phases, each as long as its slowest warp (dependent L2 gathers along one ray), not a throughput problem.")

if m2:
    w("\n## Where a `merged2` frame goes\n")
    pm = m2["roofline"]["phase_ms_per_frame"]
    r1pm = (r1m or {}).get("roofline", {}).get("phase_ms_per_frame", {})
    descm = {"classify+start_set": "`k_classify`", "fixpoint|bundling": "bundle sort (CUB pairs), `k_bundle_heads/merge`, `k_bord_hash`, `k_bundle_order` (all rehash phases of the reference's `unordered_map` in ONE cluster launch), `k_bundle_scan`, `k_bundle_loglik`",
             "ray_emit": "`k_emit_merged` (records laid out by (bundle rank, step))", "record_sort": "CUB `DeviceRadixSort::SortKeys` on the voxel bits only: 4 stable passes (round 1: 7-8 over all bits)",
             "alloc+tile_heads": "`k_block_init`, `k_voxel_heads` (segments -> long / short queues)", "tile_apply": "`k_voxel_apply_long<1,DEEP>` (the hot voxels' chains, one warp each) || `k_voxel_apply_long` (warp per role of a voxel with >= 256 records) || `k_voxel_apply_short_t` (thread per voxel) on three streams", "frame": ""}
    w("| phase | ms (r02) | ms (r01) | what runs (r02) |")
    w("|---|---|---|---|")
    for k in ("classify+start_set", "fixpoint|bundling", "ray_emit", "record_sort", "alloc+tile_heads", "tile_apply", "frame"):
        w(f"| {k} | {f(pm.get(k), '{:.3f}')} | {f(r1pm.get(k), '{:.3f}')} | {descm[k]} |")
    sh, tot = launch_shares(os.path.join(R2, "launches_merged2_21.csv"))
    src_l = "r02/launches_merged2_21.csv"
    if not sh:
        sh, tot = launch_shares(os.path.join(R2, "launches_merged2_17.csv")); src_l = "r02/launches_merged2_17.csv"
    if sh:
        w(f"\nncu launch list `{src_l}` (serialised, so the two apply kernels ADD here while they overlap in the bench): " + ", ".join(f"`{k}` {p:.0f} %" for k, p, _ in sh[:7]) + ".")
    for tag, label in (("prof_apply_merged2_21", "the three update kernels of the final commit"), ("prof_apply_merged2_17", "the three update kernels before the hot-voxel pipelines were unrolled (deep instance 1.53 ms)"), ("prof_apply_merged2_12", "apply kernels two commits earlier (no deep instance, thread-per-voxel short kernel at 2 CTAs/SM)"), ("prof_apply_merged2", "apply kernels of the previous commit (warp-per-voxel short kernel)"), ("prof_sort_merged2", "the four record-sort passes")):
        nr = ncu_rows(os.path.join(R2, tag + ".raw.csv"))
        if nr:
            w(f"\n`ncu --set full`, {label} — `r02/{tag}.details.txt`:")
            for d in nr:
                nm = str(d.get("Kernel Name", "?"))[:48]
                w(f"* `{nm}`: {f(d.get('gpu__time_duration.sum'), '{:.3f}')} ms, DRAM {f((d.get('dram__bytes_read.sum', 0) + d.get('dram__bytes_write.sum', 0)) / 1e6, '{:.0f}')} MB, "
                  f"{f(d.get('launch__registers_per_thread'), '{:.0f}')} registers, achieved occupancy {f(d.get('sm__warps_active.avg.pct_of_peak_sustained_active'))} %, "
                  f"IPC {f(d.get('sm__inst_executed.avg.per_cycle_active'), '{:.2f}')}, SM busy {f(d.get('sm__throughput.avg.pct_of_peak_sustained_elapsed'))} %")

w("\n## Roofline (HBM) of the dominant kernel and of the frame\n")
w("`achieved` = algorithmic bytes of one frame (`U·(34+8C) + 5·P`, SURVEY.md §8d) ÷ duration; peak = `MEASURED_PEAKS.json` copy bandwidth.\n")
w("| workload | algorithmic bytes / frame | dominant phase (kernel) | ms | frac | whole frame ms | frame frac | ncu DRAM traffic of that phase |")
w("|---|---|---|---|---|---|---|---|")
for name, d in (("fast5", fin), ("merged2", m2)):
    if not d:
        continue
    r = d["roofline"]
    w(f"| {name} | {f(r['algorithmic_bytes_per_launch'] / 1e6)} MB | {r['phase']} ({r['kernel'][:60]}) | {f(r['kernel_ms'], '{:.3f}')} | {f(r['frac'], '{:.3f}')} | {f(r['frame_ms'], '{:.3f}')} | {f(r['frame_frac'], '{:.4f}')} | "
      f"{f((r.get('traffic') or 0) / 1e6)} MB (`{r.get('traffic_source')}`) |")
w("\nReading: neither workload is bandwidth bound.  `fast5` moves ≈15 MB per frame (2 µs of HBM time) through a 0.4 ms chain of dependent grid-wide phases;")
w("`merged2`'s update kernels run sequential per-voxel recurrences (they must, for bit-identical results) and are bound by instruction issue and L2 latency of a few thousand warps.")

w("\n## Tuning sweeps and experiments that lost (all on the B200 box, `bench.py --quick`)\n")
for fn in ("tuning_10.log", "tuning_12.log", "tuning_13.log", "tuning_16.log", "tuning_17.log", "tuning_18.log", "tuning_19.log", "tuning_20.log", "tuning_21.log", "tuning_22.log"):
    p = os.path.join(R2, fn)
    if os.path.exists(p):
        w(f"`r02/{fn}`:\n```")
        w(open(p).read().strip())
        w("```")
w("""
* rank groups for the observed-set solver (finish a prefix of the rays, then the rest): every extra group costs more grid barriers than it saves work -> ONE group is the default (2047 -> 2494 frames/s).
* `__noinline__` helpers / rolled loops in the solve kernel (instruction-cache theory): 20 % slower (`r02/bench_full_9.json`) - reverted.
* one CTA per hot voxel with a producer / consumer shared-memory ring (`k_voxel_apply_hot`, `KSG_HOT_KERNEL=1`) and a warp-wide ray walk for the `merged` emit (`KSG_EMIT_WARP=1`): 100 and 132 frames/s against 148 (`r02/tuning_10.log`) - kept as opt-in.
* exact parallel scan of the hot voxels' float chains (`hot_voxel_mode` 1 / 2, `csrc/ksg_hot.cuh`, parity green): slower than the per-voxel kernels it feeds (`r02/bench_merged2_hot*.json`) - off by default.
* what DID pay for `merged`, in order: per-voxel work items instead of per-tile CTAs (62 -> 141 frames/s), capping the short kernel's residency so that the long one runs beside it (148 -> 166), the thread-per-voxel short kernel (-> 178), a separate deep-pipeline instance for the hot voxels' chains (-> 194), the bare-addition weight chain (-> 200), loop unrolling instead of register rotation in those pipelines and running the small long-segment kernel behind the short one (-> 209), and issuing a batch's 32 shuffles ahead of the weight chain (final number above).  The lesson of the last three: a software pipeline written as `a = b; b = load()` waits for the load it has just issued, and a chain whose every link waits for its own shuffle runs at the shuffle's latency, not the adder's.
* an L2 persistence window (`cudaAccessPropertyPersisting`) for the `L·freq` rows (`KSG_L2_PERSIST=1`): no effect (217.2 vs 216.5 frames/s, `r02/tuning_22.log`) - the rows are L2 resident anyway.
* thread-per-voxel kernel for the short segments: 20x fewer warp instructions than the warp-per-voxel kernel (ncu: 1.6 G -> see above) but no faster standalone (dependent gathers at low occupancy); it wins by leaving the SMs to the long-segment kernel (1 CTA/SM: 178 frames/s against 170).
""")

w("## Multi-GPU (one process per GPU, NCCL; device-timed, max over ranks)\n")
rows = []
for n in (2, 4, 8):
    for tag, what in (("bench_seq_fast5_n%d.json", "`fast5`, one sequence + map per GPU (replicas, no collective)"),
                      ("bench_spatial_merged2_n%d.json", "`merged2`, ONE map sharded by tile owner, frame broadcast with NCCL (strong scaling)"),
                      ("bench_frames_fast5_n%d.json", "`fast5`, frame-per-GPU batches: NCCL all-gather of the delta maps + merge"),
                      ("bench_frames_720p_c150_n%d.json", "configs[3] 1280x720 / 5 cm / C = 150, frame-per-GPU batches"),
                      ("bench_spatial_merged1_4k_c40_n%d.json", "configs[4] 3840x2160 / 1 cm / C = 40 `merged`, spatially sharded")):
        d = load(os.path.join(R2, tag % n))
        if d:
            extra = ""
            if d.get("collective"):
                extra = f"{d['collective']['bytes_per_step'] / 1e6:.0f} MB all-gathered per batch; integrate {d['phase_ms_per_step']['integrate_own_frame']:.2f} + gather {d['phase_ms_per_step']['all_gather']:.2f} + merge {d['phase_ms_per_step']['merge_all_deltas']:.2f} ms"
            elif d.get("roofline"):
                p = d["roofline"]["phase_ms_per_frame"]
                extra = f"apply {p.get('tile_apply', 0):.2f} of {p.get('frame', 0):.2f} ms per frame"
            rows.append(f"| {n} | {what} | {f(d['value'])} | {f(d['e2e']['value'])} | {d.get('scaling')} | {extra} | `r02/{tag % n}` |")
            for k, v in (d.get("workloads") or {}).items():      # the shared-map modes measured inside the driver's own N-GPU command line
                if not v or "value" not in v:
                    continue
                if v.get("collective"):
                    ex = f"{v['collective']['bytes_per_step'] / 1e6:.0f} MB all-gathered per batch; integrate {v['phase_ms_per_step']['integrate_own_frame']:.2f} + gather {v['phase_ms_per_step']['all_gather']:.2f} + merge {v['phase_ms_per_step']['merge_all_deltas']:.2f} ms"
                else:
                    pp = v["roofline"]["phase_ms_per_frame"]
                    ex = f"apply {pp.get('tile_apply', 0):.2f} of {pp.get('frame', 0):.2f} ms per frame"
                rows.append(f"| {n} | `workloads.{k}` of that line | {f(v['value'])} | {f(v['e2e']['value'])} | {v.get('scaling')} | {ex} | `r02/{tag % n}` |")
for tag, what in (("bench_fast5_720p_c150_n1.json", "configs[3] geometry on ONE GPU (sequential)"), ("bench_merged1_4k_c40_n1.json", "configs[4] on ONE GPU")):
    d = load(os.path.join(R2, tag))
    if d:
        p = d["roofline"]["phase_ms_per_frame"]
        rows.append(f"| 1 | {what} | {f(d['value'], '{:.2f}')} | {f(d['e2e']['value'], '{:.2f}')} | — | {f(d['mvoxel_updates_per_s'], '{:.0f}')} Mvoxel-updates/s, frame {p.get('frame', 0):.2f} ms | `r02/{tag}` |")
if rows:
    w("| GPUs | mode | frames/s (resident) | frames/s (e2e) | scaling | notes | file |")
    w("|---|---|---|---|---|---|---|")
    out.extend(rows)
w("\nReading: replicas scale with the number of sequences: the N = 2 row is the final commit (same synthetic stream on every rank: 2 x 2470 = 1.00 x the 1-GPU rate); the N = 8 row was taken earlier with one trajectory per rank, where the slowest stream sets the time (0.92).  The frame-per-GPU mode, with voxel-granular deltas, is 2.2x (`fast5`) to 2.9x (configs[3]) faster on 8 GPUs than ONE GPU running the frames sequentially - its own-frame integration carries the clearing of the delta layers and one host synchronisation per batch, and gather + merge grow with N.  Spatial sharding divides only the per-voxel update: `merged2` 175 -> 244 frames/s on 8 GPUs (ray casting, bundling and the record sort are replicated on every rank: Amdahl), configs[4] is sort / emit bound at N = 8 (its 2 timed steps also allocate thousands of new 725 KB blocks, which is why `value` is below the profiled frame time).")
w("\nReal-NCCL parity: `tests/test_gpu_multi.py` (the sharded map assembled from the ranks' exports equals the unsharded map, bit for bit) — `r02/gpu_multi_n*.log`.")

w("\n## GPU test logs\n")
for fn in sorted(glob.glob(os.path.join(R2, "gpu_suite_*.log")) + glob.glob(os.path.join(R2, "gpu_quick_*.log"))):
    last = [l for l in open(fn).read().splitlines() if "passed" in l or "failed" in l]
    w(f"* `r02/{os.path.basename(fn)}`: {last[-1].strip() if last else '(see file)'}")
w("\nOlder `r02/bench_*` files (`*_v2`, `*_s3`, `*_7`, `*_8`, `*_9`, `merged*_hot*` ...) are the measurements of intermediate commits of this round, kept for the history of each decision; `r02/ubench_launch.txt` = launch / graph / grid-barrier / read-back latencies measured on the box.")
open(os.path.join(ROOT, "profiles", "README.md"), "w").write("\n".join(out) + "\n")
print("ok", fin_src)
