// hzb_search.cuh -- the reference's horizon search as a per-lane state machine.
//
// Host/device source: the CUDA kernels (horizon.cu) compile it for the device; the CPU test
// infrastructure compiles THE SAME SOURCE for the host and drives it with CPU ray casts
// (tests/test_oracle_cpu.py::test_product_state_machine_on_the_cpu), so that the search logic --
// including the packet companions and the cast accounting -- is covered by the CPU suite.
// Nothing here touches a BVH: the state machine only says which table index to cast next.
#pragma once
#include "hzb_hd.cuh"
#include "hzb_tri.cuh"

namespace hzb {

struct SearchTables {   // trig tables and limits of one call (horizon_comp.cpp:711-731)
    const float* __restrict__ azim_sin; const float* __restrict__ azim_cos;
    const float* __restrict__ elev_ang; const float* __restrict__ elev_sin; const float* __restrict__ elev_cos;
    int azim_num, elev_num;
    float acc, low, up, dist; double step;
};

struct Frame {  // per-cell local frame: columns east, north, norm (horizon_comp.cpp:773-779)
    F3 org;
    float m00, m01, m02, m10, m11, m12, m20, m21, m22;
};

HZB_HD Frame make_frame(F3 vert, F3 norm, F3 north, float lift) {
    Frame f;
    f.org = f3(__fadd_rn(vert.x, __fmul_rn(norm.x, lift)), __fadd_rn(vert.y, __fmul_rn(norm.y, lift)),
               __fadd_rn(vert.z, __fmul_rn(norm.z, lift)));
    const float ex = __fsub_rn(__fmul_rn(north.y, norm.z), __fmul_rn(north.z, norm.y));
    const float ey = __fsub_rn(__fmul_rn(north.z, norm.x), __fmul_rn(north.x, norm.z));
    const float ez = __fsub_rn(__fmul_rn(north.x, norm.y), __fmul_rn(north.y, norm.x));
    f.m00 = ex; f.m01 = north.x; f.m02 = norm.x;
    f.m10 = ey; f.m11 = north.y; f.m12 = norm.y;
    f.m20 = ez; f.m21 = north.z; f.m22 = norm.z;
    return f;
}

HZB_HD F3 ray_dir(const SearchTables& s, const Frame& f, int ie, int k) {
    const float ec = __ldg(s.elev_cos + ie), es = __ldg(s.elev_sin + ie);
    const float r0 = __fmul_rn(ec, __ldg(s.azim_sin + k)), r1 = __fmul_rn(ec, __ldg(s.azim_cos + k)), r2 = es;
    return f3(__fadd_rn(__fadd_rn(__fmul_rn(f.m00, r0), __fmul_rn(f.m01, r1)), __fmul_rn(f.m02, r2)),
              __fadd_rn(__fadd_rn(__fmul_rn(f.m10, r0), __fmul_rn(f.m11, r1)), __fmul_rn(f.m12, r2)),
              __fadd_rn(__fadd_rn(__fmul_rn(f.m20, r0), __fmul_rn(f.m21, r1)), __fmul_rn(f.m22, r2)));
}


HZB_HD int index_of(const SearchTables& s, float elev) {  // (int)roundf((elev-low)/(acc/5.0))
    const double q = __ddiv_rn((double)__fsub_rn(elev, s.low), s.step);
    return (int)roundf(__double2float_rn(q));
}
HZB_HD float midpoint(float a, float b) { return __fmul_rn(__fadd_rn(a, b), 0.5f); }


// ===========================================================================
// Search state machine.  Every lane owns one cell and runs the reference's
// search (horizon_comp.cpp:302-498) as explicit states, so that all lanes of a
// warp can share ONE traversal loop (hzb_wq.cuh) instead of sitting in three
// inlined copies of it: sm_advance consumes the result of the last cast and
// returns the table index of the next cast, or "cell finished".  The hit/miss
// DECISIONS are those of cell_search<ALG> above, bit for bit.
// ===========================================================================
struct LaneSM {
    // search state (horizon_comp.cpp:387-498 unrolled into states)
    int phase;       // 0 idle/no cell, 1 bisect, 2 upward, 3 downward, 4 discrete; 5 / 6: prelude of an azimuth segment (below)
    int k, cur, prev, count, prev_az;   // during a bisection (phase 1) prev / count hold the bits of lim_up / lim_low
    int spec_ie; bool spec_hit;   // packet kernels: table index / result of the cast that travelled with the last one (-1: none)
};

// Azimuth segments (tail of a launch, horizon.cu).  The guess_constant chain hands the previous azimuth's table
// index to the next one (:439-441), so a lane that starts in the middle of a cell's chain -- at azimuth k_start
// -- needs the index the chain holds there.  Every step of the chain moves the index by a multiple of 10 (casts
// at prev+5+10m, result in the middle of the last hit and the first miss), so the index keeps the residue the
// first azimuth's bisection gave it, and the chain's value at k_start-1 is the rung of that residue class
// between the highest hit and the lowest miss.  The PRELUDE finds it without the chain: phase 5 repeats the
// bisection of azimuth 0 (no output) for the residue, phase 6 bisects the rungs {c0 + 10 j} at azimuth
// k_start-1.  The result is a GUESS -- hit(elevation) need not be monotone, the table ends clamp -- and is
// reported through seg_guess; the caller's fix-up pass compares it with the index the preceding segment really
// produced and recomputes the segment if they differ, so the outputs are the sequential chain's in every case.
// Prelude casts are not reference casts (the caller does not count them).

template <int ALG, bool PK>
HZB_HD bool sm_begin_azimuth(const SearchTables& s, LaneSM& m, int& cast_ie, int& lo_ie, bool prelude = false) {
    // returns true if a cast is required (cast_ie set), false if the azimuth needs none
    const int top = s.elev_num - 1;
    if (ALG == 0) {
        m.phase = 4; m.prev = 0; m.cur = min(10, top); cast_ie = m.cur;
        if (PK) lo_ie = min(m.cur + 10, top);                   // the next sample, should this one hit
        return true;
    } else if (ALG == 1 || m.k == 0) {
        m.phase = (ALG == 2 && prelude) ? 5 : 1; m.prev = __float_as_int(s.up); m.count = __float_as_int(s.low);
        m.cur = index_of(s, midpoint(s.up, s.low));
        const float ea = __ldg(s.elev_ang + m.cur);
        if (fmaxf(__fsub_rn(s.up, ea), __fsub_rn(ea, s.low)) > s.acc) { cast_ie = m.cur; return true; }
        return false;
    } else {
        m.phase = 2; m.count = 0;
        m.prev = max(m.prev_az - 5, 0); m.cur = min(m.prev + 10, top); cast_ie = m.cur;
        if (PK) lo_ie = max(min(m.prev_az + 5, top) - 10, 0);   // first index of the downward search (:472-476)
        return true;
    }
}

// Consume the result of the last cast (if any) and move on until the next cast
// is known or the cell is finished.  Returns true with cast_ie set when a ray
// must be traced; false when the cell is complete.
// PK (packet kernels): every cast of the stepping searches names a COMPANION in lo_ie --
// the cast the reference makes next if this one goes the expected way (prev-5 beside the
// first prev+5 of a guess_constant azimuth, otherwise the next index in the direction of
// travel).  The kernel traces both as one packet and records the companion's index and
// result in m.spec_ie / m.spec_hit; when the search then asks for exactly that index the
// stored result is consumed instead of casting, and counted in extra_rays -- i.e. only
// when the reference would have cast it.  An unused companion result is dropped.
// The search ends in front of azimuth k_end (< 0: azim_num).  k_start > 0 (guess_constant only): the lane owns the
// azimuths [k_start, k_end) of its cell and starts with m.k == 0, m.phase == 0: prelude first (see above);
// seg_guess receives the chain index the prelude found.  r0_known >= 0: the index the bisection of azimuth 0 ended
// with is already known (the lane that owns the head of the chain reports it through seg_guess, k_start == 0), so
// the prelude skips its first step.
template <int ALG, bool PK, typename OB>
HZB_HD bool sm_advance(const SearchTables& s, LaneSM& m, bool have_result, bool hit, OB& ob, int& cast_ie,
                                           int& lo_ie, unsigned int& extra_rays, const int k_start = 0, int k_end = -1,
                                           int* seg_guess = nullptr, const int r0_known = -1) {
    const int top = s.elev_num - 1;
    if (k_end < 0) k_end = s.azim_num;
    lo_ie = -1;
#define HZB_SM_CAST(COMPANION)                                                                        \
    do {                                                                                              \
        if (PK && m.spec_ie == m.cur) { m.spec_ie = -1; hit = m.spec_hit; ++extra_rays; goto again; } \
        cast_ie = m.cur;                                                                              \
        if (PK) { lo_ie = (COMPANION); if (lo_ie == cast_ie) lo_ie = -1; }                            \
        return true;                                                                                  \
    } while (0)
again:
    while (true) {
        if (!have_result) {  // start of an azimuth
            if (ALG == 2 && k_start > 0 && m.k == 0 && r0_known >= 0) { m.phase = 5; m.cur = r0_known; have_result = true; goto bisect_done; }
            if (sm_begin_azimuth<ALG, PK>(s, m, cast_ie, lo_ie, k_start > 0)) { if (lo_ie == cast_ie) lo_ie = -1; return true; }
            // bisect needed no cast at all: fall through to "azimuth finished" with phase 1
            hit = false; have_result = true;
            // (emulate loop exit below)
            goto bisect_done;
        }
        if (m.phase == 1 || (ALG == 2 && m.phase == 5)) {
            {
                const float ea = __ldg(s.elev_ang + m.cur);
                if (hit) m.count = __float_as_int(ea); else m.prev = __float_as_int(ea);
                const float lim_up = __int_as_float(m.prev), lim_low = __int_as_float(m.count);
                m.cur = index_of(s, midpoint(lim_up, lim_low));
                const float ea2 = __ldg(s.elev_ang + m.cur);
                if (fmaxf(__fsub_rn(lim_up, ea2), __fsub_rn(ea2, lim_low)) > s.acc) { cast_ie = m.cur; return true; }
            }
        bisect_done:
            if (ALG == 2 && m.phase == 5) {
                // prelude, second step: rungs c0 + 10 j of the chain's residue class at azimuth k_start-1;
                // j = m.prev counts as a hit, j = m.count as a miss (virtual rungs below the table / at its top)
                m.prev_az = (m.cur + 5) % 10;
                m.prev = -1; m.count = (top - m.prev_az) / 10 + 1;
                m.phase = 6; m.k = k_start - 1;
                goto ladder_next;
            }
            ob.put(m.k, midpoint(__int_as_float(m.prev), __int_as_float(m.count)));   // un-quantised midpoint (:377, :428)
            m.prev_az = m.cur;            // seeds the chain (:429)
            if (ALG == 2 && seg_guess) *seg_guess = m.cur;     // (the head of a split chain tells the other segments)
        } else if (ALG == 2 && m.phase == 6) {
            if (m.cur >= top) hit = false;            // termination rules of the chain
            if (m.cur == 0) hit = true;
            if (hit) m.prev = (m.cur - m.prev_az) / 10; else m.count = (m.cur - m.prev_az) / 10;
        ladder_next:
            if (m.count - m.prev > 1) {
                m.cur = m.prev_az + 10 * ((m.prev + m.count) >> 1);
                cast_ie = m.cur; lo_ie = -1;
                return true;
            }
            {   // the chain's own arithmetic on the bracketing rungs (clamped like :443-447, :472-476)
                const int lo_i = max(m.prev_az + 10 * m.prev, 0), hi_i = min(m.prev_az + 10 * m.prev + 10, top);
                m.prev_az = index_of(s, midpoint(__ldg(s.elev_ang + lo_i), __ldg(s.elev_ang + hi_i)));
                if (seg_guess) *seg_guess = m.prev_az;
            }
            // m.k == k_start-1: "azimuth finished" below moves on to the segment's first azimuth
        } else if (m.phase == 2) {
            m.count++;
            if (m.cur == top) hit = false;            // termination rule
            if (hit) { m.prev = m.cur; m.cur = min(m.cur + 10, top); HZB_SM_CAST(min(m.cur + 10, top)); }
            if (m.count <= 1) {                       // first upward cast missed: search downwards (:471-488)
                m.phase = 3;
                m.prev = min(m.prev_az + 5, top); m.cur = max(m.prev - 10, 0);
                HZB_SM_CAST(max(m.cur - 10, 0));
            }
            const int ie = index_of(s, midpoint(__ldg(s.elev_ang + m.prev), __ldg(s.elev_ang + m.cur)));
            ob.put_idx(m.k, ie, __ldg(s.elev_ang + ie)); m.prev_az = ie;      // a table entry: index and value
        } else if (m.phase == 3) {
            if (m.cur == 0) hit = true;               // termination rule
            if (!hit) { m.prev = m.cur; m.cur = max(m.cur - 10, 0); HZB_SM_CAST(max(m.cur - 10, 0)); }
            const int ie = index_of(s, midpoint(__ldg(s.elev_ang + m.prev), __ldg(s.elev_ang + m.cur)));
            ob.put_idx(m.k, ie, __ldg(s.elev_ang + ie)); m.prev_az = ie;
        } else {  // phase 4: discrete sampling (:309-331)
            if (m.cur == top) hit = false;
            if (hit) { m.prev = m.cur; m.cur = min(m.cur + 10, top); HZB_SM_CAST(min(m.cur + 10, top)); }
            ob.put(m.k, midpoint(__ldg(s.elev_ang + m.prev), __ldg(s.elev_ang + m.cur)));
        }
        // azimuth finished
        m.k++; m.spec_ie = -1;
        if (m.k >= k_end) { m.phase = 0; return false; }
        have_result = false;
    }
#undef HZB_SM_CAST
}

}  // namespace hzb
