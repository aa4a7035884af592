"""Summarise an ncu raw + source CSV pair for the tcgen05 kernel (per-tile numbers)."""
import csv, sys
from collections import defaultdict, Counter
raw, src = sys.argv[1], sys.argv[2]
tiles = float(sys.argv[3]) if len(sys.argv) > 3 else 148 * 886
rows = list(csv.reader(open(raw))); hdr = rows[0]
for k in ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
          'smsp__inst_executed.sum','smsp__cycles_active.avg','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
          'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__issue_active.avg.pct_of_peak_sustained_active',
          'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
          'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','launch__registers_per_thread','sm__warps_active.avg.pct_of_peak_sustained_active']:
    if k in hdr: print(f"{k:72s} {rows[2][hdr.index(k)]} {rows[1][hdr.index(k)]}")
rows = list(csv.reader(open(src))); hdr = rows[1]; data = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
def f(r, k):
    try: return float(r[ix[k]])
    except Exception: return 0.0
tot_s = sum(f(r, '# Samples') for r in data); tot_i = sum(f(r, 'Instructions Executed') for r in data)
print('inst/tile', round(tot_i / tiles, 1), 'smem wavefronts/tile', round(sum(f(r, 'L1 Wavefronts Shared') for r in data) / tiles, 1),
      'ideal', round(sum(f(r, 'L1 Wavefronts Shared Ideal') for r in data) / tiles, 1))
op_i = defaultdict(float); op_s = defaultdict(float)
for r in data:
    t = r[ix['Source']].strip().split()
    op = (t[0] if not t[0].startswith('@') else t[1]).split('.')[0]
    op_i[op] += f(r, 'Instructions Executed'); op_s[op] += f(r, '# Samples')
print(' '.join(f"{op}:{op_i[op]/tiles:.0f}" for op in sorted(op_i, key=lambda o: -op_i[o])[:26]))
stalls = [h for h in hdr if h.startswith('stall_') and 'Not' not in h]
print(' '.join(f"{s_[6:]}:{sum(f(r,s_) for r in data)/tot_s*100:.1f}" for s_ in sorted(stalls, key=lambda s: -sum(f(r, s) for r in data))[:10]))
W = 48
for i in range(0, len(data), W):
    blk = data[i:i + W]
    smp = sum(f(r, '# Samples') for r in blk); ex = sum(f(r, 'Instructions Executed') for r in blk) / tiles
    if smp / tot_s < 0.004: continue
    ops = Counter()
    for r in blk:
        t = r[ix['Source']].strip().split()
        op = (t[0] if not t[0].startswith('@') else t[1]).split('.')[0]
        if op in ('LDTM','SYNCS','UTCHMMA','UTMALDG','UTMASTG','LDS','STS','F2FP','DADD','ATOMS','STG','SHFL','BAR','LDG','NANOSLEEP','UTCBAR','MUFU','VOTE','LDL','STL','FMNMX3'):
            ops[op] += 1
    st = {s_[6:]: sum(f(r, s_) for r in blk) for s_ in stalls}
    top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    print(f"{blk[0][ix['Address']][-5:]} smp {smp/tot_s*100:5.1f}% exec/tile {ex:7.1f} {dict(ops)} {[(a, round(b/max(smp,1)*100)) for a,b in top]}")
