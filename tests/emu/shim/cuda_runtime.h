/* empty: under QZ_WARP_EMU the kernel sources see tests/emu/warp_emu.h instead of the CUDA runtime (test infrastructure) */
