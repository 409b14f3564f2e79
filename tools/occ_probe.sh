for pad in 0 60; do echo "pad=$pad KB"; CPPFLOW_ASM_SMEM_PAD_KB=$pad python tools/probe_kernels.py 2>&1 | grep -E "^all assemble|^flags|^pose"; done
