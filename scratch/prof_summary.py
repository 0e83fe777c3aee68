"""Turn gpurun_out/launches_<tag>.csv and gpurun_out/prof_<tag>.ncu-rep into profiles/<tag>_*.md + traffic.json."""
import csv, io, json, re, subprocess, sys, collections
tag = sys.argv[1]
# ---- launch list ------------------------------------------------------------------------------------
rows = [r for r in csv.reader(l for l in open('gpurun_out/launches_%s.csv' % tag) if l.startswith('"'))]
hdr = rows[0]; ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
agg = collections.OrderedDict()
for r in rows[1:]:
    name = re.sub(r'\(.*', '', r[ki]).replace('void ', '')
    v = float(r[vi].replace(',', '')); u = r[ui]
    us = v / 1000.0 if u in ('ns', 'nsecond') else (v if u in ('us', 'usecond') else v * 1000.0)
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += us
tot = sum(a[1] for a in agg.values())
with open('profiles/%s_launch_shares_1M.md' % tag, 'w') as f:
    f.write('# %s: ncu launch list of `python bench.py --steps 2 --warmup 3 --no-cpu-baseline` (DFSPH, 1 M particles, fast kernels)\n\n' % tag)
    f.write('`ncu --metrics gpu__time_duration.sum --clock-control none -c 900` (cold-cache, serialised: compare SHARES). Raw list: %s_launches_bench_1M.csv\n\n' % tag)
    f.write('| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|\n')
    for n, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write('| %s | %d | %.1f | %.1f | %.1f %% |\n' % (n[:90], c, us, us / c, 100 * us / tot))
subprocess.run(['cp', 'gpurun_out/launches_%s.csv' % tag, 'profiles/%s_launches_bench_1M.csv' % tag])
# ---- full capture ------------------------------------------------------------------------------------
out = subprocess.run(['ncu', '-i', 'gpurun_out/prof_%s.ncu-rep' % tag, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out))); hdr, units = rows[0], rows[1]
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio']
seen = collections.OrderedDict()
for r in rows[2:]:
    n = re.sub(r'\(.*', '', r[hdr.index('Kernel Name')]).replace('void ', '')
    seen[n] = r      # keep the last (warm) launch of each kernel
traffic = {}
with open('profiles/%s_ncu_dfsph_1M.md' % tag, 'w') as f:
    f.write('# %s: ncu --set full, DFSPH 1 M particles, fast kernels (current build)\n\n' % tag)
    f.write("`ncu --set full --clock-control none --import-source on -k regex:'k_df_|k_build_lists' -s 68 -c 12 python scratch/t_perf1.py 100 2`\n\n")
    f.write('| metric | ' + ' | '.join(seen) + ' |\n|---|' + '---|' * len(seen) + '\n')
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            f.write('| %s (%s) | ' % (k, units[i]) + ' | '.join(r[i] for r in seen.values()) + ' |\n')
    for n, r in seen.items():
        def val(k):
            i = hdr.index(k); v = float(r[i].replace(',', '')); u = units[i]
            return v * {'Mbyte': 1e6, 'Gbyte': 1e9, 'Kbyte': 1e3, 'byte': 1}.get(u, 1)
        short = {'k_df_drho': 'df_drho', 'k_df_div_iter': 'df_div_iter', 'k_df_ext_force': 'df_ext_force', 'k_df_rho_adv': 'df_rho_adv',
                 'k_df_vel_adv_iter': 'df_vel_adv_iter', 'k_df_warm_start': 'df_warm_start', 'k_build_lists': 'lists', 'k_df_position': 'df_position'}
        for a, b in short.items():
            if a in n: traffic[b] = val('dram__bytes_read.sum') + val('dram__bytes_write.sum')
    f.write('\nDRAM traffic per launch (read + write), used as `roofline.traffic` by bench.py: ' + json.dumps({k: round(v / 1e6, 1) for k, v in traffic.items()}) + ' MB\n')
old = json.load(open('profiles/traffic.json')); old.update(traffic)
json.dump(old, open('profiles/traffic.json', 'w'), indent=1)
print(open('profiles/%s_launch_shares_1M.md' % tag).read()[:2500]); print(open('profiles/%s_ncu_dfsph_1M.md' % tag).read())
