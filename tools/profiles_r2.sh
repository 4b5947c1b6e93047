#!/bin/bash
# here (no GPU): turns gpurun_out/<tag>* of tools/final_r2.sh into the tracked summaries under profiles/
tag=${1:-r2g2}
P=profiles
cp gpurun_out/${tag}_bench_default.json $P/r2_bench_default.json
cp gpurun_out/${tag}_bench_reference.json $P/r2_bench_reference.json
[ -e gpurun_out/${tag}_bench_2gpu.json ] && cp gpurun_out/${tag}_bench_2gpu.json $P/r2_bench_2gpu.json; [ -e gpurun_out/${tag}_bench_8gpu.json ] && cp gpurun_out/${tag}_bench_8gpu.json $P/r2_bench_8gpu.json; true
cp gpurun_out/${tag}_launches.csv $P/r2_launches.csv
cp gpurun_out/${tag}_launches_gzip1m.csv $P/r2_launches_gzip1m.csv
{
  echo "# ncu --set full --clock-control none, config 2 (4096 x 64 KiB zlib), one launch of each hot kernel (gpurun_out/${tag}.ncu-rep)"
  python tools/ncusum.py ${tag} 2>&1
  echo
  echo "# stall reasons per issued instruction, pipe utilisation"
  ncu -i gpurun_out/${tag}.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]
ks=[k for k in h if 'issue_stalled' in k and 'per_issue_active' in k and 'not_issued' not in k]
for k in ks+['sm__cycles_active.avg','sm__cycles_elapsed.max','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','l1tex__throughput.avg.pct_of_peak_sustained_active','sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.sum']:
    if k in h:
        i=h.index(k); print(k.replace('smsp__average_warps_issue_stalled_','stall ').replace('_per_issue_active.ratio',''), ' | '.join(r[i][:12] for r in rows[2:]))
"
} > $P/r2_ncu_summary.txt
{
  echo "# k_inflate_resolve (inflate_copy.cuh): hottest lines"
  python profiles/hotlines.py gpurun_out/${tag}.ncu-rep 3bz_b200/libthreebz_cuda.so k_inflate_resolve 28 2>&1 | cut -c1-200
  echo
  echo "# k_inflate_decode (huff_decode.cuh): hottest lines"
  python profiles/hotlines.py gpurun_out/${tag}.ncu-rep 3bz_b200/libthreebz_cuda.so k_inflate_decode 32 2>&1 | cut -c1-200
} > $P/r2_hotlines.txt
{
  echo "# cuobjdump -sass 3bz_b200/libthreebz_cuda.so, k_inflate_decode: the input of every decode lane arrives through cp.async"
  echo "# (LDGSTS.E.BYPASS.128 = cp.async.cg.shared.global 16 bytes; LDGDEPBAR = commit_group; DEPBAR.LE SB0 = wait_group)"
  cuobjdump -sass 3bz_b200/libthreebz_cuda.so 2>/dev/null | awk '/Function : /{f=(index($0,"k_inflate_decodePK")>0)} f' | grep -E "LDGSTS|LDGDEPBAR|DEPBAR" | sed 's/ *\/\* 0x[0-9a-f]* \*\///'
  echo
  echo "# the refill of the decode loop (huff_decode.cuh: 'entering a chunk'):"
  cuobjdump -sass 3bz_b200/libthreebz_cuda.so 2>/dev/null | awk '/Function : /{f=(index($0,"k_inflate_decodePK")>0)} f' | grep -B14 -A6 "LDGSTS" | sed 's/ *\/\* 0x[0-9a-f]* \*\///' | sed -n '60,110p'
  echo
  echo "# instruction mix of the library's kernels (no UTMA*/UBLKCP: nothing here moves tiles; LDGSTS is the one async copy)"
  for k in k_inflate_decodePK k_inflate_resolvePK k_member_crc k_inflate_seq; do
    printf "%-22s " $k; cuobjdump -sass 3bz_b200/libthreebz_cuda.so 2>/dev/null | awk -v k=$k '/Function : /{f=(index($0,k)>0)} f' | grep -oE "^\s+/\*[0-9a-f]+\*/\s+(@!?U?P[0-9T] )?[A-Z0-9_.]+" | awk '{print $NF}' | sed 's/\..*//' | sort | uniq -c | sort -rn | head -8 | awk '{printf "%s %s  ", $2, $1}'; echo
  done
} > $P/r2_sass_ldgsts.txt
{
  echo "# same-box A/B runs of round 2 (tools/sweep.sh: every library variant on the bench workload, per-kernel CUDA-event times)"
  for f in r2n r2p r2q r2r r2s r2x; do [ -e gpurun_out/${f}_sweep.log ] && { echo "## $f"; grep -E "^==|^\[tbz\]" gpurun_out/${f}_sweep.log; }; done
  echo "## phase two, third session of round 2 (variants of inflate_copy.cuh; var_rN = N token rounds over the pending queue before pointer jumping,"
  echo "## nopj = token rounds only, 47 KB and 4 CTAs/SM; w3072t3 = 3 072-byte windows of 768 tokens, 4 CTAs/SM; tNNN = NNN threads per CTA)"
  echo "## r2j: the round variants ran at 2 CTAs/SM (77 KB of shared memory: one CTA fewer) - kept as the occupancy data point"
  for f in r2j r2k2 r2l r2z; do [ -e gpurun_out/${f}_sweep.log ] && { echo "## $f"; grep -E "^==|^\[tbz\]" gpurun_out/${f}_sweep.log; }; done
  echo "## r2m: e2e with two small leading parts (TBZ_PIPE_FIRST = divisor; 0 = equal parts), split decode stages after the translate / CRC changes"
  cat gpurun_out/r2m_e2e.log gpurun_out/r2m_split.log 2>/dev/null
  echo "## r2i: split decode stages before (1 GiB gzip member)"; grep "tbz split" gpurun_out/r2i_gzip1g.err 2>/dev/null | tail -7
  echo "## r2t2: Kraft early exit in the block-start validation; 16 Ki / 8 Ki-symbol ring with far sources from global memory (not kept); chunk sweep"
  cat gpurun_out/r2t2_stages.log gpurun_out/r2t2_split.log 2>/dev/null
  echo "## r2u: chunk-size sweep, 1 GiB and 256 MiB members"; cat gpurun_out/r2u_split.log 2>/dev/null
  echo "## r2v2: block-start search with lane-parallel validation (same chunks found)"; cat gpurun_out/r2v2_stages.log 2>/dev/null
  echo "## r2w2: stage times per chunk size"; cat gpurun_out/r2w2_stages.log 2>/dev/null
  echo "## r2x2: + first-filter compaction; the slowest chunks of the decode (a chunk behind a false-positive start decodes twice as far)"; cat gpurun_out/r2x2_stages.log 2>/dev/null
  echo "## r2fp: the false positives of the block-start search (same bit positions at every chunk size)"; cat gpurun_out/r2fp.log 2>/dev/null
  echo "## r2fp2: BFINAL = 0 candidates only, stricter distance-code rule: no false positive left, times smooth in the chunk size"; cat gpurun_out/r2fp2.log 2>/dev/null
  echo "## r2ds: sub-chunk size of the split decode's phase one (default 4000 bits)"; cat gpurun_out/r2ds.log 2>/dev/null
  echo "## r2pipe: e2e pipeline, result records by kernel stores (default) against a DMA; streams; the host timeline"; cat gpurun_out/r2pipe.log 2>/dev/null
  echo "## r2rv / r2rv2 / r2rv3: symbolic resolve, threads per CTA (res512 = 512 x 2 tokens became the default after r2rv)"; cat gpurun_out/r2rv.log gpurun_out/r2rv2.log gpurun_out/r2rv3.log 2>/dev/null
  echo "## r2aa: e2e pipeline geometry (parts x streams -> GB/s, ms per step, ceiling)"; grep "^parts" /tmp/r2aa.out 2>/dev/null
} > $P/r2_experiments.txt
{
  echo "# compute-sanitizer over the parity tests (config2_subset, config4_members, edge_mix [, truncation, corruption])"
  echo "## racecheck --racecheck-report analysis: the only lines reported are the acquire / release accesses of pointer jumping (inflate_copy.cuh step 4; DESIGN.md 9)"
  grep -E "Race reported|RACECHECK SUMMARY|passed" gpurun_out/r2ab_racecheck.txt | sed 's/=========//' | cut -c1-170 | sort | uniq -c | sort -rn
  echo "## memcheck"; tail -3 gpurun_out/r2ab_memcheck.txt
  echo "## memcheck over the third session's code (tools/experiments/memcheck_r2.py: split decode of a gzip and a zlib member, CRC kernel on mixed lengths, pipelined batch of 2 304 members)"; tail -3 gpurun_out/r2_memcheck2.txt
} > $P/r2_sanitizer.txt
ls -la $P | grep r2_
