// emu_primitives.h -- included by csrc/directsum.cuh INSIDE namespace lpm when LPM_CUDA_EMU is defined:
// CPU stand-ins for the inline-PTX primitives (mbarrier, TMA bulk copy, MUFU.RCP64H, LOP3).
// A "TMA copy" is a memcpy by the issuing thread followed by a phase flip of the barrier word.
inline void mbar_init(uint64_t* bar, uint32_t) { __atomic_store_n(bar, (uint64_t)0, __ATOMIC_RELEASE); }
inline void mbar_fence_init() {}
inline void mbar_expect_tx(uint64_t*, uint32_t) {}
inline bool mbar_try_wait(uint64_t* bar, uint32_t parity)
{
    const uint64_t done = __atomic_load_n(bar, __ATOMIC_ACQUIRE);     // completed phases
    if ((done & 1) != parity) return true;
    sched_yield();
    return false;
}
inline void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar)
{
    std::memcpy(dst_smem, src_gmem, bytes);
    __atomic_fetch_add(bar, (uint64_t)1, __ATOMIC_RELEASE);
}
// MUFU.RCP64H returns a reciprocal good to ~20 bits with a zero low word: truncate the exact one
inline double rcp_approx_f64(double d)
{
    const double r = 1.0 / d;
    return __hiloint2double(__double2hiint(r), 0);
}
inline int lop3_and_or(int hi, int half) { return (hi & (int)0xfffff000) | half; }
