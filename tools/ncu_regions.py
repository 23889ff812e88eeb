"""Per-code-region stall/opcode breakdown of one kernel from an .ncu-rep source page.
usage: python tools/ncu_regions.py report.ncu-rep [n_segments]"""
import csv, subprocess, sys, collections
rep = sys.argv[1]; nseg = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
nxt = [i for i, r in enumerate(rows) if i > 1 and r and r[0] == 'Kernel Name']      # several launches: the first one
data = [r for r in rows[2:(nxt[0] if nxt else None)] if len(r) == len(hdr)]
ix = {h: i for i, h in enumerate(hdr)}
def iv(r, k):
    try: return int(r[ix[k]] or 0)
    except ValueError: return 0
tot = sum(iv(r, '# Samples') for r in data)
print("kernel:", rows[0][1][:80], "instructions:", len(data), "samples:", tot)
n = len(data); seg = max(1, n // nseg)
ST = ['stall_long_sb', 'stall_short_sb', 'stall_wait', 'stall_selected', 'stall_math', 'stall_lg', 'stall_mio', 'stall_no_inst', 'stall_barrier', 'stall_membar', 'stall_sleep']
for s0 in range(0, n, seg):
    ch = data[s0:s0 + seg]
    sm = sum(iv(r, '# Samples') for r in ch); ie = sum(iv(r, 'Instructions Executed') for r in ch)
    st = {k[6:]: sum(iv(r, k) for r in ch) for k in ST}
    st = {k: v for k, v in st.items() if v}
    print(f"{s0:5d} samples {sm:5d} inst {ie:9d}", st)
print("top instructions by samples")
for r in sorted(data, key=lambda r: -iv(r, '# Samples'))[:30]:
    print(data.index(r), iv(r, '# Samples'), r[ix['Source']][:90], {k[6:]: iv(r, k) for k in ST if iv(r, k)})
