mkdir -p gpurun_out
(CE_ATTN_FIXED=1 timeout 120 python scripts/attn_timing.py; cd scripts/micro && nvcc -gencode arch=compute_100a,code=sm_100a -o mma_rate mma_rate.cu -lcuda 2>&1 | tail -2; ./mma_rate) 2>&1 | tee gpurun_out/attn_mma_stamps_r2d.log
