// tests/hostcheck/warp_emu.h -- TEST TOOLING ONLY.  Host emulation of the CUDA warp collectives used by the product's
// warp-cooperative device code (prt_b200/csrc/entry_list.cuh), so that the CPU test-suite can run that code unmodified:
// one std::thread per lane, every collective is a rendezvous of the 32 lanes.  Only what entry_list.cuh uses is provided.
#pragma once
#include <algorithm>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <cuda_runtime.h>      // float4, uint2, make_float4 ... (plain C++ types on the host)

namespace warp_emu {
struct State {
    std::barrier<> bar{32};
    uint64_t slot[32];
};
inline State *g_state = nullptr;
inline thread_local int t_lane = 0;
inline void sync() { g_state->bar.arrive_and_wait(); }
template <class T> inline T exchange(T v, int src, bool keep_own) {
    static_assert(sizeof(T) <= 8, "32- or 64-bit values only");
    uint64_t u = 0; std::memcpy(&u, &v, sizeof(T));
    g_state->slot[t_lane] = u;
    sync();
    const uint64_t r = keep_own ? u : g_state->slot[src & 31];
    sync();
    T out; std::memcpy(&out, &r, sizeof(T));
    return out;
}
}  // namespace warp_emu

inline void __syncwarp(unsigned = 0xFFFFFFFFu) { warp_emu::sync(); }
inline unsigned __ballot_sync(unsigned, bool p) {
    warp_emu::g_state->slot[warp_emu::t_lane] = p ? 1u : 0u;
    warp_emu::sync();
    unsigned r = 0u;
    for (int i = 0; i < 32; i++) r |= (unsigned)(warp_emu::g_state->slot[i] & 1u) << i;
    warp_emu::sync();
    return r;
}
inline bool __any_sync(unsigned m, bool p) { return __ballot_sync(m, p) != 0u; }
template <class T> inline T __shfl_sync(unsigned, T v, int src) { return warp_emu::exchange(v, src, false); }
template <class T> inline T __shfl_xor_sync(unsigned, T v, int o) { return warp_emu::exchange(v, warp_emu::t_lane ^ o, false); }
template <class T> inline T __shfl_up_sync(unsigned, T v, int o) { return warp_emu::exchange(v, warp_emu::t_lane - o, warp_emu::t_lane - o < 0); }
inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline int __ffs(int x) { return __builtin_ffs(x); }
inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
inline uint32_t __float_as_uint(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
inline float __uint_as_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
template <class T> inline T __ldg(const T *p) { return *p; }
inline uint32_t atomicOr(uint32_t *p, uint32_t v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
inline int atomicAdd(int *p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline int atomicSub(int *p, int v) { return __atomic_fetch_sub(p, v, __ATOMIC_RELAXED); }
// position of the offset-th set bit of mask counted from bit `base` (PTX fns.b32); 0xFFFFFFFF when there is none
inline unsigned __fns(unsigned mask, unsigned base, int offset) {
    if (offset == 0) return ((mask >> base) & 1u) ? base : 0xFFFFFFFFu;
    if (offset > 0) {
        for (unsigned b = base; b < 32u; b++) if (((mask >> b) & 1u) && --offset == 0) return b;
        return 0xFFFFFFFFu;
    }
    for (int b = (int)base; b >= 0; b--) if (((mask >> b) & 1u) && ++offset == 0) return (unsigned)b;
    return 0xFFFFFFFFu;
}
inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline unsigned long long atomicMin(unsigned long long *p, unsigned long long v) {
    unsigned long long old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (v < old && !__atomic_compare_exchange_n(p, &old, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}
namespace prt { using std::max; using std::min; }
