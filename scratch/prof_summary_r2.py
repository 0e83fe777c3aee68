"""gpurun_out/launches_<TAG>.csv + gpurun_out/prof_<TAG>_<group>.ncu-rep  ->  profiles/<TAG>_launch_shares_1M.md,
profiles/<TAG>_ncu_<group>.md and profiles/traffic.json (DRAM bytes per launch of every profiled kernel).
usage: python scratch/prof_summary_r2.py TAG [directory of the .ncu-rep files] [output directory]"""
import collections, csv, glob, io, json, os, re, subprocess, sys
tag = sys.argv[1]
REP = sys.argv[2] if len(sys.argv) > 2 else 'gpurun_out'
OUT = sys.argv[3] if len(sys.argv) > 3 else 'profiles'
SHORT = {'k_df_drho': 'df_drho', 'k_df_div_iter': 'df_div_iter', 'k_df_ext_force': 'df_ext_force', 'k_df_rho_adv': 'df_rho_adv',
         'k_df_vel_adv_iter': 'df_vel_adv_iter', 'k_df_warm_start': 'df_warm_start', 'k_build_lists': 'lists', 'k_df_position': 'df_position',
         'k_pc_ext_force': 'pc_ext', 'k_pc_predict_rho': 'pc_rho', 'k_pc_press_force': 'pc_force', 'k_ii_advect': 'ii_adv',
         'k_ii_rho_adv_aii': 'ii_aii', 'k_ii_dij': 'ii_dij', 'k_ii_update_p': 'ii_update', 'k_wc_force': 'wc_force'}
def clean(n): return re.sub(r'\(.*', '', n).replace('void ', '').replace('sph_fast::', '').replace('sph_strict::', 'strict::')
# ---- launch list ------------------------------------------------------------------------------------
lp = 'gpurun_out/launches_%s.csv' % tag
if os.path.exists(lp):
    rows = [r for r in csv.reader(l for l in open(lp) if l.startswith('"'))]
    hdr = rows[0]; ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    agg = collections.OrderedDict()
    for r in rows[1:]:
        v = float(r[vi].replace(',', '')); u = r[ui]
        us = v / 1000.0 if u in ('ns', 'nsecond') else (v if u in ('us', 'usecond') else v * 1000.0)
        a = agg.setdefault(clean(r[ki]), [0, 0.0]); a[0] += 1; a[1] += us
    tot = sum(a[1] for a in agg.values())
    with open(OUT + '/%s_launch_shares_1M.md' % tag, 'w') as f:
        f.write('# %s: ncu launch list of `python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-also` (DFSPH, 10^6 particles, fast kernels)\n\n' % tag)
        f.write('`ncu --metrics gpu__time_duration.sum --clock-control none -c 1200` (cold-cache, serialised: compare SHARES, not absolute times). Raw list: %s_launches_bench_1M.csv\n\n' % tag)
        f.write('| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|\n')
        for n, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write('| %s | %d | %.1f | %.1f | %.1f %% |\n' % (n[:90], c, us, us / c, 100 * us / tot))
    subprocess.run(['cp', lp, OUT + '/%s_launches_bench_1M.csv' % tag])
# ---- full captures ------------------------------------------------------------------------------------
KEYS = [('gpu__time_duration.sum', 'duration'), ('dram__bytes_read.sum', 'DRAM read'), ('dram__bytes_write.sum', 'DRAM write'),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'DRAM throughput % of peak'),
        ('sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'FMA (FP32) pipe % active'),
        ('sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'FMA pipe instr % of peak'),
        ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue slots % busy'),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps active % of max'),
        ('launch__registers_per_thread', 'registers / thread'),
        ('l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'L1 data-pipe wavefronts % of peak'),
        ('l1tex__t_sector_hit_rate.pct', 'L1 hit rate %'), ('lts__t_sector_hit_rate.pct', 'L2 hit rate %'),
        ('smsp__thread_inst_executed_per_inst_executed.ratio', 'active lanes per instruction'),
        ('smsp__inst_executed.sum', 'warp instructions'),
        ('smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'long-scoreboard stalls per issue')]
tpath = OUT + '/traffic.json'
traffic = json.load(open('profiles/traffic.json')) if os.path.exists('profiles/traffic.json') else {}
traffic.setdefault('kernels', {})
for rep in sorted(glob.glob(REP + '/prof_%s_*.ncu-rep' % tag)):
    group = re.sub(r'.*prof_%s_(.*)\.ncu-rep' % tag, r'\1', rep)
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    if len(rows) < 3: continue
    hdr, units = rows[0], rows[1]
    ti = hdr.index('gpu__time_duration.sum')
    best = collections.OrderedDict()      # per kernel: the longest launch (gated no-op launches of a finished loop are ~3 us)
    for r in rows[2:]:
        n = clean(r[hdr.index('Kernel Name')])
        if n not in best or float(r[ti].replace(',', '')) > float(best[n][ti].replace(',', '')): best[n] = r
    log = 'gpurun_out/ncu_full_%s_%s.log' % (tag, group)
    run = [l for l in open(log).read().splitlines() if ' N ' in l and 'iters' in l] if os.path.exists(log) else []
    with open(OUT + '/%s_ncu_%s.md' % (tag, group), 'w') as f:
        f.write('# %s: `ncu --set full --clock-control none --profile-from-start off`, group `%s`\n\n' % (tag, group))
        f.write('One step after the warm-up (scratch/prof_r2.sh, scratch/t_prof_solver.py); fast kernels; per kernel the longest captured launch.\n')
        if run: f.write('Run: `%s` (solver, block, N, iterations div / den / pcisph / iisph of the profiled step, error flags)\n' % run[-1])
        f.write('\n| metric | ' + ' | '.join(best) + ' |\n|---|' + '---|' * len(best) + '\n')
        for k, label in KEYS:
            if k in hdr:
                i = hdr.index(k)
                f.write('| %s (%s) | ' % (label, units[i]) + ' | '.join(r[i] for r in best.values()) + ' |\n')
        f.write('\n')
    for n, r in best.items():
        def val(k):
            i = hdr.index(k); v = float(r[i].replace(',', '')); u = units[i]
            return v * {'Mbyte': 1e6, 'Gbyte': 1e9, 'Kbyte': 1e3, 'byte': 1}.get(u, 1)
        for a, b in SHORT.items():
            if n.startswith(a) and group != 'wcsph' or (group == 'wcsph' and a == 'k_wc_force' and n.startswith(a)):
                npart = 1000000
                traffic['kernels'][b] = {'dram_bytes_per_launch': val('dram__bytes_read.sum') + val('dram__bytes_write.sum'),
                                         'particles': npart, 'source': 'profiles/%s_ncu_%s.md' % (tag, group)}
traffic['source'] = 'ncu --set full captures of round %s (per kernel: see "source"); dram__bytes_read.sum + dram__bytes_write.sum per launch' % tag
json.dump(traffic, open(tpath, 'w'), indent=1)
for p in sorted(glob.glob(OUT + '/%s_*.md' % tag)): print(p)
