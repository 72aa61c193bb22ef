"""Turn the ncu artefacts of a GPU call (gpurun_out/<call>/) into the committed summaries under profiles/.

    python tools/make_profile_summaries.py <call-dir> <tag>        e.g.  gpurun_out/c7 r2

Reads (whichever exist): launches.csv (ncu --metrics gpu__time_duration.sum launch list of bench.py), fused_step.ncu-rep and
gemm.ncu-rep (ncu --set full captures).  Writes profiles/<tag>_launches.md / .csv, <tag>_fused_step_ncu.md, <tag>_gemm_ncu.md
and profiles/fused_step_traffic.json (DRAM bytes of the forward t >= 2 launch: bench.py's `roofline.traffic`)."""
import collections, csv, io, json, os, re, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
os.makedirs(OUT, exist_ok=True)
call = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "c1")
tag = sys.argv[2] if len(sys.argv) > 2 else "r2"


def launches(csv_path, md_path, note):
    rows = [r for r in csv.reader(open(csv_path)) if len(r) > 5]
    hdr = rows[0]
    idx = {h: i for i, h in enumerate(hdr)}
    agg = collections.OrderedDict()
    tot, n = 0.0, 0
    for r in rows[1:]:
        if r[idx["Metric Name"]] != "gpu__time_duration.sum":
            continue
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "")[:86]
        v = float(r[idx["Metric Value"]].replace(",", ""))
        u = r[idx["Metric Unit"]]
        v = v / 1000 if u in ("ns", "nsecond") else (v * 1000 if u in ("ms", "msecond") else v)
        agg.setdefault(name, [0, 0.0])
        agg[name][0] += 1
        agg[name][1] += v
        tot += v
        n += 1
    with open(md_path, "w") as f:
        f.write(f"# ncu launch list ({tag}) — `gpu__time_duration.sum`, `--clock-control none`\n\n{note}\n\n")
        f.write(f"{n} launches captured, {tot:.0f} us of kernel time (cold-cache, serialised: compare SHARES).\n\n")
        f.write("| share | total us | launches | kernel |\n|---:|---:|---:|---|\n")
        for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| {100 * v / tot:.1f} % | {v:.1f} | {c} | `{k}` |\n")


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active"]


def raw_rows(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(raw)))


def full(rep, md_path, title, extra=""):
    rows = raw_rows(rep)
    hdr = rows[0]
    with open(md_path, "w") as f:
        f.write(f"# {title}\n\n`ncu --set full --clock-control none --import-source on` (values per launch).{extra}\n\n")
        for r in rows[2:]:
            f.write(f"## `{r[hdr.index('Kernel Name')][:150]}`\n\n| metric | value |\n|---|---|\n")
            for w in WANT:
                if w in hdr:
                    i = hdr.index(w)
                    f.write(f"| `{w}` | {r[i]} {rows[1][i]} |\n")
            f.write("\n")
        src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
        secs, cur = [], None
        for r in csv.reader(io.StringIO(src)):
            if r and r[0] == "Kernel Name":
                cur = {"name": r[1], "rows": []}
                secs.append(cur)
            elif cur is not None:
                cur["rows"].append(r)
        for sec in secs:
            if not sec["rows"]:
                continue
            h = sec["rows"][0]
            ix = {x: i for i, x in enumerate(h)}
            data = [r for r in sec["rows"][1:] if len(r) > 10]
            if "# Samples" not in ix or not data:
                continue
            si, ei = ix["# Samples"], ix["Instructions Executed"]
            tot = sum(int(r[si]) for r in data)
            f.write(f"### hottest SASS lines by warp-stall samples ({tot} samples) of `{sec['name'][:100]}`\n\n| samples | executed | SASS |\n|---:|---:|---|\n")
            for i in sorted(range(len(data)), key=lambda i: -int(data[i][si]))[:14]:
                f.write(f"| {data[i][si]} | {data[i][ei]} | `{data[i][ix['Source']].strip()[:90]}` |\n")
            f.write("\n")


def traffic(rep, edges, atoms):
    rows = raw_rows(rep)
    hdr = rows[0]

    def num(r, name):
        i = hdr.index(name)
        v = float(r[i].replace(",", ""))
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(rows[1][i], 1)

    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        m = re.search(r"k_bond_step_fused<([^>]*)>", name)
        if not m:
            continue
        targs = [int(re.sub(r"[^0-9]", "", a.replace("true", "1").replace("false", "0")) or 0) for a in m.group(1).split(",")]
        if len(targs) >= 4 and targs[1] == 0 and targs[3] == 0:      # <ACT, FIRST = 0, HAS_BIAS, MODE = 0, ...>: forward, t >= 2
            rd, wr = num(r, "dram__bytes_read.sum"), num(r, "dram__bytes_write.sum")
            json.dump({"kernel": "k_bond_step_fused<RELU, t>=2, forward>", "directed_edges": edges, "atoms": atoms, "precision": "bf16",
                       "dram_bytes_read": int(rd), "dram_bytes_write": int(wr), "dram_bytes_per_launch": int(rd + wr),
                       "algorithmic_bytes": 3 * edges * 300 * 2 + 12 * edges + 4 * atoms,
                       "source": f"profiles/{tag}_fused_step_ncu.md (ncu --set full, one launch of bench.py's C2 batch)"},
                      open(os.path.join(OUT, "fused_step_traffic.json"), "w"), indent=1)
            print("traffic:", rd + wr)
            return
    print("WARNING: forward t >= 2 launch not found")


p = os.path.join(call, "launches.csv")
if os.path.exists(p):
    shutil.copy(p, os.path.join(OUT, f"{tag}_launches.csv"))
    launches(p, os.path.join(OUT, f"{tag}_launches.md"),
             "Command: `ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv python bench.py --steps 2 --warmup 3 "
             "--no-cpu --no-graph` (C2: bf16 tier, 10 k molecules; warm-up + timed steps of the eager and loader-driven loops).")
p = os.path.join(call, "fused_step.ncu-rep")
if os.path.exists(p):
    full(p, os.path.join(OUT, f"{tag}_fused_step_ncu.md"), f"Fused depth-step kernel ({tag})",
         " Template arguments <ACT, FIRST, HAS_BIAS, MODE, FAR, DROP>: launches in order = forward first step (`FIRST`, reads H_0 only, "
         "also stores M^1), forward t >= 2 step (the roofline kernel), backward mirror step with tau' mask + G output (MODE 1), last "
         "backward mirror step with the tau'(H_0) mask (MODE 3).")
    b = os.path.join(call, "bench.json")
    E, V = 502100, 249437
    if os.path.exists(b):
        try:
            d = json.load(open(b))["details"]
            E, V = d["directed_edges_per_batch"], d["atoms_per_batch"]
        except Exception:
            pass
    traffic(p, E, V)
p = os.path.join(call, "gemm.ncu-rep")
if os.path.exists(p):
    full(p, os.path.join(OUT, f"{tag}_gemm_ncu.md"), f"tcgen05 linear / weight-gradient kernels ({tag})")
