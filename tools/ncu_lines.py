"""Per-source-line instruction and stall-sample shares of one kernel from an ncu report (read on the CPU box).

    python tools/ncu_lines.py report.ncu-rep k_deposit_tile4 [top]
"""
import csv, subprocess, sys, io

def load(rep, kern):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "-k",
                          "regex:" + kern], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    cur, hdr, out = None, None, []
    def num(s):
        try: return int(s)
        except Exception: return 0
    for r in rows:
        if not r: continue
        if r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
        if r[0] == 'Line No': hdr = r; continue
        if hdr and r[0].isdigit():
            d = dict(zip(hdr, r))
            out.append((cur, int(r[0]), r[1], num(d.get('Instructions Executed')), num(d.get('# Samples')), d))
    return out

if __name__ == "__main__":
    rep, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 60
    out = load(rep, kern)
    tot = sum(o[3] for o in out) or 1; tots = sum(o[4] for o in out) or 1
    print('total warp instructions', tot, 'samples', tots)
    for o in sorted(out, key=lambda x: -x[3])[:top]:
        print(f"{o[0][:20]:20s} {o[1]:4d} {100*o[3]/tot:5.1f}% s{100*o[4]/tots:5.1f}%  {o[2][:120]}")
