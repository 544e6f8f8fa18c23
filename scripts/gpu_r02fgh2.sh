# z-chunk count of the fused FGH + divergence sweep at 255^3
bash scripts/gpu_ab.sh r02fgh2 nscube255 "FDMB_FGH_CHUNKS=16" "FDMB_FGH_CHUNKS=32" "FDMB_FGH_CHUNKS=64" "FDMB_FGH_CHUNKS=128" 2>&1 | grep -E "==|steps/s|ns_fgh" | sed 's/ e2e=.*//'
