// hzb_queue.cuh -- work queue of a horizon launch: band first, interior in row order, the last tiles as azimuth segments
// (horizon.cu, "Queue order and azimuth segments").
//
// Host/device source (see hzb_hd.cuh): pure integer arithmetic on the queue fields of HorizonParams, written against
// any struct P that has those fields, so that the CPU suite can enumerate whole queues on the host
// (tests/test_oracle_cpu.py::test_queue_order_covers_every_task_once).
#pragma once
#include "hzb_hd.cuh"

namespace hzb {

// One record per (split cell, segment >= 1) and one per split cell.  guess: the chain index the segment's prelude assumed at
// the azimuth in front of the segment (guess_constant), SEG_OK where no assumption is needed, SEG_REDO after a full traversal
// stack, SEG_NONE (the buffer's initial value) if the task never ran.  casts: reference casts the task counted.
struct SegRecord { int guess; unsigned int casts; };
constexpr int SEG_NONE = -1, SEG_REDO = -2, SEG_OK = -3;
constexpr int SEG_COUNT = 4;     // segments of a split cell (the lane keeps the segment number in two spare bits of its cell word)

// Derived queue fields from (seg_count, q_by0, q_by1, q_bx, q_tail) and the launch's tile grid / block sharding.
template <typename P>
inline void queue_sections(P& p, int tiles_x, int tiles_y) {
    p.q_tiles_x = tiles_x; p.q_tiles_y = tiles_y; p.q_wi = tiles_x - 2 * p.q_bx;
    p.q_nA1 = (unsigned int)p.q_by0 * (unsigned int)tiles_x; p.q_nA2 = (unsigned int)(tiles_y - p.q_by1) * (unsigned int)tiles_x;
    p.q_nA3 = (unsigned int)(p.q_by1 - p.q_by0) * 2u * (unsigned int)p.q_bx;
    p.q_nI = (unsigned int)(p.q_by1 - p.q_by0) * (unsigned int)p.q_wi;
    p.q_total = p.q_nA1 + p.q_nA2 + p.q_nA3 + p.q_nI + p.q_tail * (unsigned int)(p.seg_count - 1);
    // first split tile (interior tile nI - tail) and the end of the interior as GLOBAL block rows of the launch's row range
    p.q_gb_end = p.q_by1 * p.blk_stride + p.blk_offset;
    p.q_gb_tail = p.q_gb_end; p.q_tx_tail = 0;
    if (p.q_tail > 0 && p.q_wi > 0) {
        const unsigned int first = p.q_nI - p.q_tail;
        p.q_gb_tail = (p.q_by0 + (int)(first / (unsigned int)p.q_wi)) * p.blk_stride + p.blk_offset;
        p.q_tx_tail = p.q_bx + (int)(first % (unsigned int)p.q_wi);
    }
}

// queue entry q -> tile (ty, tx) and task: seg 0 = the whole chain, 1 + n = azimuth segment n
template <typename P>
HZB_HD void queue_decode(const P& p, unsigned int q, int& ty, int& tx, int& seg) {
    seg = 0;
    if (q < p.q_nA1) { ty = (int)(q / p.q_tiles_x); tx = (int)(q % p.q_tiles_x); return; }
    q -= p.q_nA1;
    if (q < p.q_nA2) { ty = p.q_by1 + (int)(q / p.q_tiles_x); tx = (int)(q % p.q_tiles_x); return; }
    q -= p.q_nA2;
    if (q < p.q_nA3) {
        const int w2 = 2 * p.q_bx, c = (int)(q % w2);
        ty = p.q_by0 + (int)(q / w2); tx = c < p.q_bx ? c : p.q_tiles_x - w2 + c;
        return;
    }
    q -= p.q_nA3;
    unsigned int b = q;
    if (q >= p.q_nI - p.q_tail) {
        const unsigned int u = q - (p.q_nI - p.q_tail);
        seg = 1 + (int)(u / p.q_tail); b = p.q_nI - p.q_tail + u % p.q_tail;
    }
    ty = p.q_by0 + (int)(b / p.q_wi); tx = p.q_bx + (int)(b % p.q_wi);
}
// index of tile (local block row lb, tile column tx) among the split tiles, or -1
template <typename P>
HZB_HD int tail_tile(const P& p, int lb, int tx) {
    if (p.seg_count <= 1 || lb < p.q_by0 || lb >= p.q_by1 || tx < p.q_bx || tx >= p.q_tiles_x - p.q_bx) return -1;
    const int b = (lb - p.q_by0) * p.q_wi + (tx - p.q_bx);
    return b - (int)(p.q_nI - p.q_tail);      // < 0: interior, not split
}
// does cell (ci, cj) belong to a split tile?  (no division: compared as global block rows, host-prepared bounds)
template <typename P>
HZB_HD bool cell_is_split(const P& p, int ci, int cj) {
    const int gb = (ci - p.row_begin) >> 2, tx = cj >> 3;
    return p.seg_count > 1 && tx >= p.q_bx && tx < p.q_tiles_x - p.q_bx && gb < p.q_gb_end &&
           (gb > p.q_gb_tail || (gb == p.q_gb_tail && tx >= p.q_tx_tail));
}
template <typename P>
HZB_HD int seg_begin(const P& p, int n) { return (int)(((long long)n * p.azim_num) / SEG_COUNT); }
// the lane's cell word: row << 16 | column (both <= 32767, horizon.pyx:149-151) with the segment number in bits 15 and 31
HZB_HD int cell_row(unsigned int w) { return (int)((w >> 16) & 0x7FFFu); }
HZB_HD int cell_col(unsigned int w) { return (int)(w & 0x7FFFu); }
HZB_HD int cell_seg(unsigned int w) { return (int)(((w >> 15) & 1u) | ((w >> 30) & 2u)); }
// record of (cell of a split tile, segment n >= 1); n == SEG_COUNT: the cell's shared record (guess = the index the
// bisection of azimuth 0 ended with, published by the lane that owns segment 0)
template <typename P>
HZB_HD size_t seg_record_index(const P& p, int ci, int cj, int n) {
    const int gb = (ci - p.row_begin) >> 2, lb = (gb - p.blk_offset) / p.blk_stride;
    const int tt = tail_tile(p, lb, cj >> 3);
    const int in_tile = (((ci - p.row_begin) & 3) << 3) | (cj & 7);
    return ((size_t)tt * 32 + in_tile) * SEG_COUNT + (size_t)(n - 1);
}
// cell slots (one per task) of local block row lb: what publish_cell counts up to
template <typename P>
HZB_HD unsigned int row_slots(const P& p, int lb) {
    unsigned int n = p.row_full;
    if (p.seg_count > 1 && lb >= p.q_by0 && lb < p.q_by1) {
        const long long over = (long long)(lb - p.q_by0 + 1) * p.q_wi - (long long)(p.q_nI - p.q_tail);
        const long long split = over < 0 ? 0 : (over > p.q_wi ? p.q_wi : over);
        n += (unsigned int)split * 32u * (unsigned int)(p.seg_count - 1);
    }
    return n;
}

}  // namespace hzb
