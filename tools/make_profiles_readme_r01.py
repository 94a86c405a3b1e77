"""Regenerates the measured tables of profiles/README.md from the committed JSON / CSV files of profiles/r01."""
import csv, json, os, re
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R = os.path.join(ROOT, "profiles", "r01")
d5 = json.load(open(os.path.join(R, "bench_fast5.json"))); dr = json.load(open(os.path.join(R, "bench_fast5_reference.json")))
dm = json.load(open(os.path.join(R, "bench_merged2.json")))
SC = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "cycle": 1, "us": 1, "ms": 1e3}
def raw(path):
    rows = list(csv.reader(open(path))); hdr, units, vals = rows[0], rows[1], rows[-1]
    out = {}
    for h, u, v in zip(hdr, units, vals):
        try:
            out[h] = float(v.replace(",", "")) * SC.get(u, 1)
        except ValueError:
            pass
    return out
t5 = raw(os.path.join(R, "prof_apply_fast5.raw.csv")); tm = raw(os.path.join(R, "prof_apply_merged2.raw.csv"))
ph = d5["roofline"]["phase_ms_per_frame"]; c5 = d5["cpu_baseline"]; cm = dm["cpu_baseline"]
head = f"""# profiles/ — measured evidence (B200, sm_100a, CUDA 12.9, driver 580)

Everything here was produced on the `gpurun` B200 box; numbers taken under `ncu` are never bench values.  `r01/` = round 1
(final state of the round).  The tables below are generated from the committed files by `tools/make_profiles_readme.py`.

## r01 headline (`python bench.py`, defaults: `fast5` = 640x480 depth+label stream, 5 cm voxels, 21 classes, `fast`; {d5['steps']} steps after {d5['warmup']} warm-up)

| arm | frames/s | ms/frame | Mvoxel-updates/s | file |
|---|---|---|---|---|
| ours, frames resident in HBM (`value`, CUDA events on the launching stream) | **{d5['value']:.0f}** | {d5['ms_per_step']:.3f} | {d5['mvoxel_updates_per_s']:.0f} | `r01/bench_fast5.json` |
| ours, end to end from page-locked host frames (`e2e`: H2D of 1.5 MB/frame + counter read-backs, wall clock) | **{d5['e2e']['value']:.0f}** | {1e3/d5['e2e']['value']:.3f} | — | `r01/bench_fast5.json` |
| reference arm `bench.py --impl reference`: CPU port of the reference, fastest of 1/4/16/128 threads (= {dr['cpu_baseline']['cores']}) | {dr['value']:.1f} | {dr['ms_per_step']:.1f} | {dr['mvoxel_updates_per_s']:.1f} | `r01/bench_fast5_reference.json` |
| `cpu_baseline` inside our run (thread calibration fps: {', '.join(f'{k}: {v:.1f}' for k, v in c5['thread_calibration_fps'].items())}) | {c5['value']:.1f} | {1e3/c5['value']:.1f} | {c5['mvoxel_updates_per_s']:.1f} | `r01/bench_fast5.json` |

SM clock {d5['clocks']['sm_mhz']:.0f} MHz (= max) during the timed region, no throttle reasons.  End-to-end speed-up over the CPU
reference arm on the same box ≈ {d5['e2e']['value']/dr['value']:.0f}x.  The reference's multi-threaded mode *loses* on a 128-core host (per-voxel
mutexes, two shared atomic hash sets, thread creation per frame): 128 threads reach {c5['thread_calibration_fps'].get('128', 0):.1f} fps.

`merged2` (BASELINE configs[2]: 640x480, 2 cm, 21 classes, `merged`; ≈31.8 M voxel updates per frame):
{dm['value']:.1f} frames/s = {dm['mvoxel_updates_per_s']:.0f} Mvoxel-updates/s (e2e {dm['e2e']['value']:.1f}) vs {cm['value']:.2f} frames/s = {cm['mvoxel_updates_per_s']:.0f}
Mvoxel-updates/s for the CPU port at its best thread count ({cm['cores']}) — `r01/bench_merged2.json`.

Multi-GPU (`torchrun --nproc-per-node N bench.py --gpus N`, one stream + map per rank, no data-path collective, max over ranks):
N = 2: 2916 frames/s (1458 per GPU); N = 4: 5871 frames/s resident / 5532 end to end (1468 per GPU) — ≈97–98 % of N x the 1-GPU
rate measured in the same sessions (40 steps each).

## Where a `fast5` frame goes (CUDA events inside the library, `roofline.phase_ms_per_frame`)

| phase | ms | kernels |
|---|---|---|
| classify + start set + ray setup | {ph['classify+start_set']:.3f} | `k_depth_flags`, CUB select, `k_classify`, `k_start_push/eval/commit`, CUB select, `k_ray_setup`, 3 memsets |
| observed-set fixpoint | {ph['fixpoint|bundling']:.3f} | 6–8 × `k_eval` (4 sweeps, then one per host read-back of two counters) |
| commit + ray emit | {ph['ray_emit']:.3f} | `k_obs_commit`, `k_emit_fast` |
| record sort | {ph['record_sort']:.3f} | CUB `DeviceRadixSort` (7 one-sweep passes over ≈50–80 K keys: launch bound) |
| block alloc + tile heads | {ph['alloc+tile_heads']:.3f} | `k_block_init`, `k_tile_heads` |
| tile apply | {ph['tile_apply']:.3f} | `k_tile_apply<TMA,1,fast>` |
| frame | {ph['frame']:.3f} | ({d5['gpu_launches']/d5['steps']:.0f} own kernel launches + {d5['library_calls']/d5['steps']:.0f} CUB calls per frame) |

ncu launch list of the same workload (`ncu --metrics gpu__time_duration.sum --clock-control none`, cold caches, serialised):
`r01/launches_fast5.csv`, per-kernel table in `r01/launch_summary.md` — `k_eval` 47 %, `k_tile_apply` 12 %, radix sort 9 %,
`k_ray_setup` 7 %, `k_start_eval` 6 %: the kernel shares agree with the event-timed phases (fixpoint ≈ 37–47 % of the frame).

## Roofline of the tile-apply kernel

`roofline.achieved` = algorithmic bytes per launch (`U·(34+8C) + 5·P`, SURVEY.md §8d) ÷ the kernel's event-timed duration;
peak = {d5['roofline']['peak']:.0f} GB/s (`MEASURED_PEAKS.json`, measured copy bandwidth, "of measured").

| workload | algorithmic bytes / launch | kernel ms (events) | achieved GB/s | frac of measured | ncu DRAM traffic / launch (`dram__bytes_read.sum + write.sum`) |
|---|---|---|---|---|---|
| fast5 | {d5['roofline']['algorithmic_bytes_per_launch']/1e6:.1f} MB | {d5['roofline']['kernel_ms']:.3f} | {d5['roofline']['achieved']:.0f} | {d5['roofline']['frac']:.3f} | {t5['dram__bytes_read.sum']/1e6:.1f} MB read + {t5['dram__bytes_write.sum']/1e6:.1f} MB written (`r01/prof_apply_fast5.raw.csv`) |
| merged2 | {dm['roofline']['algorithmic_bytes_per_launch']/1e9:.2f} GB | {dm['roofline']['kernel_ms']:.2f} | {dm['roofline']['achieved']:.0f} | {dm['roofline']['frac']:.3f} | {tm['dram__bytes_read.sum']/1e6:.0f} MB read + {tm['dram__bytes_write.sum']/1e6:.0f} MB written (`r01/prof_apply_merged2.raw.csv`) |
"""
path = os.path.join(ROOT, "profiles", "r01", "README.md")
old = open(path).read()
tail = old[old.index("Reading (honest):"):]
open(path, "w").write(head + "\n" + tail)
print("ok")
