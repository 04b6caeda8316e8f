// evcursor.cuh -- a lane's cursor over its voice's device events (shared by the fused kernels and the generated ones).
#pragma once
#include <stdint.h>

#include "dev.h"
#include "nodes.cuh"

#ifndef EVC_UNCOND
#define EVC_UNCOND 1
#endif
#ifndef EVC_RAW_TAIL
#define EVC_RAW_TAIL 1
#endif

namespace kgpu {

// one parameter event (16 B) through the read-only path
KN_DEV DevEvent decode_event(const uint4 q) {
    DevEvent e;
    e.frame = q.x;
    e.node = (uint16_t)(q.y & 0xFFFFu);
    e.op = (uint16_t)(q.y >> 16);
    e.reg = q.z;
    e.value = q.w;
    return e;
}
KN_DEV uint4 ldg_event_raw(const DevEvent *p) { return __ldg(reinterpret_cast<const uint4 *>(p)); }
KN_DEV DevEvent ldg_event(const DevEvent *p) { return decode_event(ldg_event_raw(p)); }

// Per-lane event cursor with the next FOUR events in registers.  e[0] is complete by construction;
// the slot freed by a pop is refilled at once, so a load has four pops (or thousands of frames) to
// land -- a scheduler that holds a single warp has nothing else to hide a load behind.  Events
// arrive by H2D copy, i.e. from DRAM: the 32-byte sectors 16 events ahead are pulled into L2 early.
struct EvCursor {
    const DevEvent *events;
    uint32_t cur, end, next_frame;
    DevEvent e0, e1, e2;
#if EVC_RAW_TAIL
    uint4 r3; // the newest entry stays as loaded: unpacking it right behind the load made the warp wait for the load there
              // (r2j profile: 1.1 % of render_sub_asr's time at the PRMT behind the refill) instead of one pop later
#else
    DevEvent e3;
#endif
    static constexpr uint32_t AHEAD = 16;
    KN_DEV void prefetch(uint32_t i) const {
        if (i < end) asm volatile("prefetch.global.L2 [%0];" ::"l"(events + i));
    }
    KN_DEV void init(const DevEvent *ev, const uint32_t *off, uint32_t v, bool on) {
        events = ev;
        cur = end = 0;
        next_frame = 0xFFFFFFFFu;
        e0 = e1 = e2 = DevEvent{};
#if EVC_RAW_TAIL
        r3 = make_uint4(0u, 0u, 0u, 0u);
#else
        e3 = DevEvent{};
#endif
        if (ev && on) {
            cur = off[v];
            end = off[v + 1];
            if (cur < end) e0 = ldg_event(events + cur);
            if (cur + 1 < end) e1 = ldg_event(events + cur + 1);
            if (cur + 2 < end) e2 = ldg_event(events + cur + 2);
#if EVC_RAW_TAIL
            if (cur + 3 < end) r3 = ldg_event_raw(events + cur + 3);
#else
            if (cur + 3 < end) e3 = ldg_event(events + cur + 3);
#endif
            for (uint32_t i = 4; i < AHEAD; i += 2) prefetch(cur + i);
            if (cur < end) next_frame = e0.frame;
        }
    }
    KN_DEV void pop() {
        cur++;
        e0 = e1;
        e1 = e2;
#if EVC_RAW_TAIL
        e2 = decode_event(r3);
#else
        e2 = e3;
#endif
        next_frame = cur < end ? e0.frame : 0xFFFFFFFFu;
        // unconditional (index clamped to the list): a predicated load comes with a predicated move of its result, and that
        // move made the warp wait for the load right here (r2a profile: 140 cycles per pop) instead of three pops later
#if EVC_RAW_TAIL
        r3 = ldg_event_raw(events + min(cur + 3u, end - 1u));
#elif EVC_UNCOND
        e3 = ldg_event(events + min(cur + 3u, end - 1u));
#else
        if (cur + 3 < end) e3 = ldg_event(events + cur + 3);
#endif
        prefetch(cur + AHEAD);
    }
};

} // namespace kgpu
