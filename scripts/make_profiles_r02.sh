#!/bin/bash
# summaries of the captures of scripts/profile_r02.sh (gpurun_out/r02_*) -> profiles/r02_*  (run here, after the GPU call)
S=scripts/ncu_summary.py
{
echo "# Round 2 - ncu \`--set full --clock-control none --import-source on\` of the torus sweep (B200, one GPU)"; echo
echo "Command: \`scripts/profile_r02.sh\` (bench.py --config c5 --samples-per-gpu 2980 --steps 1 --warmup 1, 130 chains x 1000 x 1000; the profiled command first ran plain and exited 0)."
echo "Launch 1 = a TWO-sweep launch (red, black, red, black on a tile of 20 rows + 14 halo rows: sweeps that emit nothing), launch 2 = emitting single sweep (1 in 5: also writes the 520 MB of finished fp32 samples and forms their statistics, EMIT = 3)."; echo
python $S gpurun_out/r02_prof_sweep.ncu-rep
echo; echo "## single-sweep launches (MCLCAR_SWEEP_DOUBLE=0): plain sweep, emitting sweep (samples + statistics)"; echo
python $S gpurun_out/r02_prof_sweep_single.ncu-rep
echo; echo "## statistics-only emission (\`keep = 0\`: EMIT = 2, sufficient statistics accumulated in the emitting sweep instead of writing samples)"; echo
python $S gpurun_out/r02_prof_sweep_stats.ncu-rep
echo; echo "## config 2 (256 x 256 torus, 500 chains, TR = 64: 4 tiles per chain)"; echo
python $S gpurun_out/r02_prof_c2_sweep.ncu-rep
} > profiles/r02_ncu_sweep.md
{ echo "# Round 2 - ncu full captures of the likelihood kernels on config 5 (stats_torus_kernel<float>: 2980 fp32 samples = 11.92 GB; k3_moments: 200 psi x 2980 samples)"; echo
python $S gpurun_out/r02_prof_stats_torus.ncu-rep; echo; python $S gpurun_out/r02_prof_k3.ncu-rep; } > profiles/r02_ncu_likelihood_c5.md
{ echo "# Round 2 - ncu full capture of gibbs_csr_kernel<BINOM> (config 3: 5000-region map, 592 chains = 4 per CTA, state and graph in shared memory, independence proposals from the Gaussian approximation of the full conditional, whole chain run of 181 sweeps in one launch)"; echo
python $S gpurun_out/r02_prof_csr_sampler.ncu-rep; } > profiles/r02_ncu_csr_sampler.md
{ echo "# Round 2 - ncu full capture of glm_terms_kernel (config 3: 10^4 x 5000 fp32 latent fields, binomial data term in fp64)"; echo
echo "## final kernel: piecewise degree-8 table of log1p(exp(-a)), all betas of a tile evaluated together (first launch: the single-beta lHZy.psi pass, two-slot build; second: an 8-beta tile)"; echo
python $S gpurun_out/r02_prof_glm_terms.ncu-rep; echo
echo "## before (start of round 2): exp / log1p / divide per (sample, site, beta), 8-beta tile, 632 CTAs"; echo
cat profiles/r02_ncu_glm_terms_before.md 2>/dev/null
echo; echo "## the reductions of the same step (k3_moments / k3_finalize)"; echo; python $S gpurun_out/r02_prof_k3_glm.ncu-rep; } > profiles/r02_ncu_glm_terms.md
{ echo "# Round 2 - ncu full capture of stats_site_major_kernel<float> (config 3: z'z, z'Wz of 10^4 site-major fields on the 5000-region map)"; echo
python $S gpurun_out/r02_prof_stats_sm.ncu-rep; } > profiles/r02_ncu_stats_site_major.md
{ echo "# Round 2 - ncu full captures of the hierarchical-CAR step (config 4: K = 400 districts, N = 2 x 10^4): gibbs_csr_kernel<GAUSS> on Q*, the K-dimensional quadratic forms, combine"; echo
python $S gpurun_out/r02_prof_hcar.ncu-rep; } > profiles/r02_ncu_hcar.md
{ echo "# Round 2 - ncu launch lists (\`--metrics gpu__time_duration.sum --clock-control none\`) of \`bench.py --config cN --steps 1 --warmup 1\`"; echo
echo "Cold-cache, serialised launches: compare SHARES with the bench's own CUDA-event spans (\`kernels\` / \`roofline.share_of_step\` in r02_bench_*.json), not absolutes."
for c in c5 c2 c3 c4; do echo; echo "## $c"; echo; python scripts/launch_list_md.py gpurun_out/r02_launches_$c.csv; done; } > profiles/r02_launches.md
for c in c2 c5 c5_statsonly c5_single reference; do cp gpurun_out/r02_bench_$c.json profiles/r02_bench_$c.json; done   # c1, c3, c4, c4p: scripts/check_r02.sh (gpurun_out/r2_bench_cN.json)
cuobjdump -sass mclcar_b200/build/sampler_torus.o 2>/dev/null | awk '/Function : .*gibbs_torus_sweep_fused.*Lb1ELi0EfLi3E/{f=1;next} f&&/Function : /{f=0} f' > /tmp/fused_sass.txt
{ echo "# SASS evidence: bulk-copy engine and mbarrier instructions per kernel (\`python scripts/sass_evidence.py\`, cuobjdump -sass of the built library)"; echo; python scripts/sass_evidence.py; echo; echo "Excerpt, \`gibbs_torus_sweep_fused<SOR, EMIT = 0, float>\` tile load:"; echo '```'; grep "SYNCS\|UBLKCP\|UBLKPF" /tmp/fused_sass.txt | sed 's/^\s*//' | cut -c1-110 | head -12; echo '```'; } > profiles/r02_sass_excerpt.md
