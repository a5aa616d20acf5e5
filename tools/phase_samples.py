"""Developer tool (CPU only): warp-state samples of k_fit_resident<tile in tensor memory> from an `ncu --set full` report, summed
per PHASE of the SGD step.

ncu's source page gives samples per SASS instruction; `nvdisasm -gi` of a cubin compiled from the same source gives, per SASS
instruction, the source line and the chain of call sites it was inlined from. The outermost line inside the kernel body names
the phase (line ranges found from marker text in csrc/fit_resident.cu); the rounds of the unrolled full-resolution loop share
their source lines and are told apart by their position relative to the early gather / P3 block in the instruction stream.

    python tools/phase_samples.py gpurun_out/r2i_fit_resident.ncu-rep > profiles/r2i_fit_resident_phase_samples.txt
"""
import collections, csv, io, os, re, subprocess, sys, tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from few_shot_seg_cwt_b200 import build as B

KERNEL = "_ZN3cwt14k_fit_residentILi512ELi1ELi512ELi20ELi5ELi60ELi60ELb0ELb1ELi0EEEv14CUtensorMap_stNS_14ResidentParamsE"
SRC = os.path.join(B.CSRC, "fit_resident.cu")


def line_of(text, nth=0, after=0):
    """1-based number of the nth line at or after ``after`` that contains ``text``."""
    hits = [i + 1 for i, l in enumerate(open(SRC).read().splitlines()) if text in l and i + 1 >= after]
    return hits[nth]


def main(rep):
    # ---- SASS -> (outermost kernel-body line) from a cubin of the current source (same code as the profiled build) ----
    with tempfile.TemporaryDirectory() as td:
        cubin = os.path.join(td, "fr.cubin")
        flags = [f for f in B.FLAGS if f not in ("-Xptxas", "-v")]
        subprocess.run([B.NVCC, *flags, "-cubin", SRC, "-o", cubin], check=True, capture_output=True)
        dis = subprocess.run(["nvdisasm", "-gi", cubin], check=True, capture_output=True, text=True).stdout.splitlines()
    start = next(i for i, l in enumerate(dis) if l.startswith(".text." + KERNEL + ":"))
    ins, cur = [], None
    for l in dis[start + 1:]:
        if l.startswith("\t.section") or l.startswith(".text."):
            break
        if "//## File" in l:
            sites = re.findall(r'"([^"]+)", line (\d+)', l)
            body = [int(n) for f, n in sites if f.endswith("fit_resident.cu")]
            cur = body[-1] if body else None                  # outermost call site inside fit_resident.cu
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip(), cur))
    # ---- samples per SASS instruction from the report ----
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], check=True, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, data = rows[1], rows[2:]
    assert len(data) == len(ins), (len(data), len(ins), "the report was not taken on this source")
    for (off, text, _), r in zip(ins[::97], data[::97]):          # spot-check that the two listings are the same code
        op = [t for t in text.split() if not t.startswith("@")][0].split(".")[0]
        assert op in r[1], (off, text, r[1])
    i_s, i_x = hdr.index("# Samples"), hdr.index("Instructions Executed")
    stall_cols = [(i, h[len("stall_"):]) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]

    # ---- phases: line ranges of the kernel body ----
    L_ep = line_of("for (int e = group; e < p.E; e += p.G) {")
    L_step = line_of("for (int t = 0; t < p.T; ++t, ++gstep) {")
    L_p1_end = line_of("t_acc[0] += n - tk0")
    L_hr_loop = line_of("for (int m = 0; m < RES_MAXTASK; ++m) {", after=L_step)
    L_early = line_of("if (m == 1) {", after=L_hr_loop)
    L_halo_wait = line_of("while (!mbar_try_wait(halo_ready", after=L_hr_loop)
    L_hr_body = line_of("unsigned d = hr_desc[m];", after=L_hr_loop)
    L_late_gather = line_of("gather_part(kSplit ? TM_EARLY : 0, NP);") - 1
    L_p3 = line_of("t_acc[2] += n - tk0")
    L_red = line_of("red_add_u64(acc_t + RES_ACC(tid)", after=L_p3)
    L_ar_end = line_of("compute_sync<CT>();", after=L_red)
    L_step_end = line_of("t_acc[3] += n - tk0")
    L_halo_warp = line_of("halo warp: fetch the ring")
    L_applier = line_of("} else if (NA > 0 && warp")

    names = ["kernel setup", "episode staging (tile -> tensor memory, max |F|, coefficient table)", "P1 (tensor-memory sweep, butterfly, combine, publish)",
             "HR round 0", "gather + P3 of the early columns (halo shadow)", "halo wait", "HR round 1", "gather of the late columns",
             "P3 of the late columns + RED", "all-reduce wait, SGD apply, Wd", "episode end (drain, result)", "halo warp"]
    seen_early = False
    agg = collections.OrderedDict((n, {"samples": 0, "inst": 0, "stalls": collections.Counter()}) for n in names)
    phases = []
    for (off, text, line), r in zip(ins, data):
        if line is None:
            ph = "kernel setup"
        elif line >= L_halo_warp:
            ph = "halo warp"
        elif line >= L_applier:
            ph = "episode end (drain, result)"
        elif line < L_ep:
            ph = "kernel setup"
        elif line < L_step:
            ph = names[1]
        elif line <= L_p1_end:
            ph = names[2]
        elif line < L_late_gather:
            if L_early <= line < L_halo_wait - 1:
                ph = names[4]; seen_early = True
            elif L_halo_wait - 4 <= line < L_hr_body:
                ph = "halo wait"
            elif line < L_hr_loop:                             # the gather lambda's definition lines never carry code of their own
                ph = names[2]
            else:
                ph = "HR round 1" if seen_early else "HR round 0"
        elif line <= L_p3:
            ph = "gather of the late columns"
        elif line <= L_red:
            ph = names[8]
        elif line <= L_ar_end:
            ph = names[9]
        elif line <= L_step_end:
            ph = names[9]
        else:
            ph = "episode end (drain, result)"
        phases.append(ph)
    # Address arithmetic that the compiler re-materialises inside the step loop carries the line of the DEFINITION it derives from
    # (tid / lane / the warp's tensor-memory window: kernel-setup lines). Inside the loop's address range such an instruction
    # belongs to the phase of the code around it: it inherits the phase of the nearest earlier instruction that has one.
    loop = [i for i, ph in enumerate(phases) if ph in names[2:10]]
    for i in range(loop[0], loop[-1] + 1):
        if phases[i] in ("kernel setup", names[1]) and i > loop[0]:
            phases[i] = phases[i - 1]
    for ph, r in zip(phases, data):
        a = agg[ph]
        a["samples"] += int(r[i_s]); a["inst"] += int(r[i_x])
        for i, s in stall_cols:
            a["stalls"][s] += int(r[i] or 0)
    total = sum(a["samples"] for a in agg.values())
    n_cta, steps = 144, 200 * 16
    print(f"k_fit_resident<tile in tensor memory>, ncu --set full warp samples of {os.path.basename(rep)} summed per phase (tools/phase_samples.py):")
    print(f"E = 64, 200 steps; share of all {total} warp samples, warp instructions per CTA and SGD step, top stall reasons\n")
    for n, a in agg.items():
        if not a["samples"] and not a["inst"]:
            continue
        tot = sum(a["stalls"].values()) or 1
        top = ", ".join(f"{k} {100 * v / tot:.0f}%" for k, v in a["stalls"].most_common(3))
        print(f"{n:72s} {100 * a['samples'] / total:5.1f} %  {a['inst'] / n_cta / steps:7.0f} warp-instr/CTA/step   {top}")


if __name__ == "__main__":
    main(sys.argv[1])
