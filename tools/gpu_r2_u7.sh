# round 2, call U7: MMA micro-benchmark -- two CTAs per SM / two accumulation chains per CTA for the context model's patterns
mkdir -p gpurun_out
for p in 9 10 14 15 16 17 19 11 18; do timeout 60 tools/ubench/mma_shapes 20000 $p 20; done > gpurun_out/r2u7_mma_chains.txt 2>&1
cut -c1-160 gpurun_out/r2u7_mma_chains.txt
