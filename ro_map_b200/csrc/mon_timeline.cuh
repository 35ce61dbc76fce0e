// mon_timeline.cuh — optional per-kernel timeline of the iteration graph (build with
// MON_EXTRA_NVCC_FLAGS=-DMON_TIMELINE; tools/timeline.py).  Every CTA's thread 0 folds %globaltimer into
// [first start, last end] of its (iteration mod MON_TL_ITERS, kind) slot, so a graph replay leaves the real overlap of
// the level-pipelined branches behind — nsys is not available on the box.  Compiled out in the product build.
#pragma once
#include <cstdint>

#define MON_TL_ITERS 64
#define MON_TL_KINDS 16
// kinds: 0 batch, 1 sample points, 2-5 encode by first level / 4, 6 fused MLP, 7-10 scatter by level group,
// 11-14 optimizer by level group, 15 optimizer (MLP weights + loss)
#define MON_TL_B 0
#define MON_TL_P 1
#define MON_TL_E 2
#define MON_TL_M 6
#define MON_TL_S 7
#define MON_TL_O 11
#define MON_TL_OMLP 15

#ifdef MON_TIMELINE
struct MonTlSlot { unsigned long long start, end; };
// one table per translation unit (no relocatable device code in this build); MON_TL_DEFINE(name) also exports the
// host accessors mon_debug_tl_<name>_{read,reset}
#define MON_TL_DEFINE(name)                                                                                         \
    static __device__ MonTlSlot mon_tl_tab[MON_TL_ITERS * MON_TL_KINDS];                                             \
    extern "C" int mon_debug_tl_##name##_read(unsigned long long* out) {                                            \
        return (int)cudaMemcpyFromSymbol(out, mon_tl_tab, sizeof(mon_tl_tab));                                       \
    }                                                                                                               \
    extern "C" int mon_debug_tl_##name##_reset(void) {                                                              \
        static MonTlSlot init[MON_TL_ITERS * MON_TL_KINDS];                                                          \
        for (auto& s : init) { s.start = ~0ull; s.end = 0ull; }                                                      \
        return (int)cudaMemcpyToSymbol(mon_tl_tab, init, sizeof(init));                                              \
    }
#define MON_TL(kind, iter)                                                                                          \
    struct MonTlScope_ {                                                                                            \
        MonTlSlot* s;                                                                                               \
        __device__ static unsigned long long now() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; } \
        __device__ MonTlScope_(MonTlSlot* p) : s(threadIdx.x == 0 ? p : nullptr) { if (s) atomicMin(&s->start, now()); } \
        __device__ ~MonTlScope_() { if (s) atomicMax(&s->end, now()); }                                             \
    } mon_tl_scope_(&mon_tl_tab[((iter) % MON_TL_ITERS) * MON_TL_KINDS + (kind)])
// a point event folded into [earliest, latest] over the CTAs (e.g. "first table slice resident", "CTA done")
#define MON_TL_MARK(kind, iter)                                                                                     \
    do {                                                                                                            \
        if (threadIdx.x == 0) {                                                                                     \
            unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                           \
            MonTlSlot* s_ = &mon_tl_tab[((iter) % MON_TL_ITERS) * MON_TL_KINDS + (kind)];                           \
            atomicMin(&s_->start, t_); atomicMax(&s_->end, t_);                                                     \
        }                                                                                                           \
    } while (0)
#else
#define MON_TL_DEFINE(name)
#define MON_TL(kind, iter) do { } while (0)
#define MON_TL_MARK(kind, iter) do { } while (0)
#endif
