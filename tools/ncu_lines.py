"""Map `ncu --page source --csv` (SASS view) stall samples of one kernel back to CUDA source lines using nvdisasm line info.
usage: python tools/ncu_lines.py <report.ncu-rep> <cubin> <mangled-kernel-substring> [top]"""
import csv
import re
import subprocess
import sys


def main(rep, cubin, kern, top=40):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    si, ii = hdr.index("# Samples"), hdr.index("Instructions Executed")
    stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    sass = rows[hi + 1:]
    dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
    # walk the function of interest: collect (line) per instruction in order
    lines, cur, infn = [], None, False
    for l in dis:
        if l.startswith(".text.") or "Function" in l and "::" in l:
            pass
        m = re.match(r"\s*\.section\s+\.text\.(\S+),", l)
        if m:
            infn = kern in m.group(1)
            continue
        if not infn:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
            lines.append(cur)
    n = min(len(lines), len(sass))
    print(f"# sass rows {len(sass)}, disasm instrs {len(lines)}")
    agg = {}
    total = 0
    for k in range(n):
        s = int(sass[k][si] or 0)
        total += s
        a = agg.setdefault(lines[k], [0, 0, {}])
        a[0] += s
        a[1] += int(sass[k][ii] or 0)
        for ci, h in stall_cols:
            v = int(sass[k][ci] or 0)
            if v:
                a[2][h] = a[2].get(h, 0) + v
    src = {}
    for key, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        if key and key[0] not in src:
            try:
                src[key[0]] = open(f"/root/repo/diffusion_model_nemo_b200/csrc/{key[0]}").read().splitlines()
            except Exception:
                src[key[0]] = []
        text = src[key[0]][key[1] - 1].strip()[:90] if key and len(src.get(key[0], [])) >= key[1] else ""
        st = ", ".join(f"{h[6:]}={c}" for h, c in sorted(v[2].items(), key=lambda x: -x[1])[:3])
        print(f"{100.0 * v[0] / max(total, 1):5.1f}%  {v[0]:6d} smp  {v[1]:8d} inst  {key}  {text}   [{st}]")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 40)
