/*
 * wb_deframe_kernel.cuh -- K2: unique-word search + packet collection over the soft-decision rows.
 *
 * Replaces the LOOK_FOR_UW / COLLECT_PACKET state machine of reference src/drs232_ldpc.c:176-216
 * (v1: 40-bit RS232-framed UW, >= 35 matches, 3230 symbols) and src/wenet_ldpc.c:171-209
 * (v2: 32-bit UW, >= 28 matches, 2584 symbols).  Semantics kept:
 *   - hard bit = symbol < 0;
 *   - the sliding window (bit_buffer) is NOT advanced while a packet is being collected and is
 *     not cleared afterwards, so matching resumes over the stream with the packet excised;
 *   - the symbol that completes the UW is not part of the packet.
 *
 * One warp per stream.  While looking, 32 symbols are examined per step (fetched 32 x WB_DEFRAME_NF at a time): a ballot gives the 32 new
 * hard bits, lane j forms the window as it stands after symbol j and scores it with one popcount;
 * the first hit (lowest lane) wins.  While collecting, nothing is touched: the packet is recorded
 * as an offset into the row and the scan jumps over it.  Packets are only located here; the symbols
 * are gathered (RS232 strip / descramble) by the decoder kernel straight from the row.
 */
#ifndef WB_DEFRAME_KERNEL_CUH
#define WB_DEFRAME_KERNEL_CUH

#include "wb_internal.h"

#ifndef WB_DEFRAME_NF
#define WB_DEFRAME_NF 16
#endif

__global__ void __launch_bounds__(128)
wb_deframe_kernel(wb_deframe_params p, wb_stream_state *state, wb_cursor *cursor, const float *sd,
                  unsigned long long sd_stride, unsigned *jobs, int job_cap, int n_streams)
{
    const int lane = threadIdx.x & 31;
    const int s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (s >= n_streams) return;
    wb_stream_state &st = state[s];
    const float *row = sd + (size_t)s * sd_stride;
    unsigned *myjobs = jobs + (size_t)s * job_cap;
    const int n_new = (int)cursor[s].n_sd;

    unsigned long long W = st.window;
    int collecting = st.collecting, ind = st.ind;
    int t = 0, nj = 0;

    if (collecting) {
        int need = p.nsym - ind;
        if (n_new >= need) {
            if (lane == 0 && nj < job_cap) myjobs[nj] = (unsigned)(WB_CARRY_CAP - ind);
            nj++;
            t = need; collecting = 0; ind = 0;
        } else {
            ind += n_new; t = n_new;
        }
    }
    /* WB_DEFRAME_NF x 32 symbols are fetched per round (that many independent loads per lane in flight: the scan is a
       chain of L2 round trips otherwise), then examined 32 at a time; a hit restarts the round at the jump target.
       The round after this one is fetched before this one is examined: a stream without packets -- the slowest kind,
       every symbol is looked at -- then never waits for memory; after a hit the fetched round is dropped. */
    float va[WB_DEFRAME_NF], vb[WB_DEFRAME_NF];
#define WB_DEFRAME_FETCH(V, T0)                                                                         \
    do {                                                                                                \
        _Pragma("unroll") for (int k = 0; k < WB_DEFRAME_NF; k++) {                                     \
            const int idx = (T0) + 32 * k + lane;                                                       \
            V[k] = (idx < n_new) ? row[WB_CARRY_CAP + idx] : 0.0f;                                      \
        }                                                                                               \
    } while (0)
    if (t < n_new) WB_DEFRAME_FETCH(va, t);
    while (t < n_new) {
        const int t_next = t + 32 * WB_DEFRAME_NF;
        if (t_next < n_new) WB_DEFRAME_FETCH(vb, t_next);
        bool jumped = false;
#pragma unroll
        for (int k = 0; k < WB_DEFRAME_NF; k++) {
            if (jumped || t >= n_new) break;
            const int idx = t + lane;
            const bool valid = idx < n_new;
            const float v = va[k];
            unsigned nb = __ballot_sync(0xffffffffu, valid && (v < 0.0f));
            /* window after symbol t+lane: symbols t..t+lane appended, newest in bit 0 */
            unsigned long long Wj = (W << (lane + 1)) | (unsigned long long)(__brev(nb) >> (31 - lane));
            int score = __popcll(~(Wj ^ p.uw) & p.uw_mask);
            unsigned hits = __ballot_sync(0xffffffffu, valid && score >= p.uw_thresh);
            if (hits) {
                int first = __ffs(hits) - 1;
                W = __shfl_sync(0xffffffffu, Wj, first);
                int start = t + first + 1;               /* first collected symbol */
                int avail = n_new - start;
                if (avail >= p.nsym) {
                    if (lane == 0 && nj < job_cap) myjobs[nj] = (unsigned)(WB_CARRY_CAP + start);
                    nj++;
                    t = start + p.nsym;
                } else {
                    collecting = 1; ind = avail; t = n_new;
                }
                jumped = true;                           /* the values fetched for this round no longer line up */
            } else {
                int nvalid = min(32, n_new - t);
                W = __shfl_sync(0xffffffffu, Wj, nvalid - 1);
                t += nvalid;
            }
        }
        if (t >= n_new) break;
        if (jumped) {
            WB_DEFRAME_FETCH(va, t);
        } else {                                         /* t == t_next: the round fetched ahead is the next one */
#pragma unroll
            for (int k = 0; k < WB_DEFRAME_NF; k++) va[k] = vb[k];
        }
    }
#undef WB_DEFRAME_FETCH
    if (lane == 0) {
        st.window = W;
        st.collecting = collecting;
        st.ind = ind;
        if (nj > job_cap) nj = job_cap;              /* cannot happen: job_cap is sized from the chunk */
        cursor[s].n_jobs = (unsigned)nj;
        cursor[s].seq0 = st.seq;
        st.seq += (unsigned)nj;
    }
}

/* After the decoder has consumed the rows: move the symbols of a half-collected packet in front of
   the next chunk (row[CARRY_CAP - ind .. CARRY_CAP)).  The collected run always ends at the end of
   the new data, and is contiguous in the row because the previous carry sits right before it. */
__global__ void __launch_bounds__(128)
wb_carry_kernel(wb_stream_state *state, const wb_cursor *cursor, float *sd, unsigned long long sd_stride,
                int n_streams)
{
    const int lane = threadIdx.x & 31;
    const int s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (s >= n_streams) return;
    const wb_stream_state &st = state[s];
    if (!st.collecting || st.ind <= 0) return;
    const int n_new = (int)cursor[s].n_sd;
    if (n_new == 0) return;
    float *row = sd + (size_t)s * sd_stride;
    const int ind = st.ind;
    const int src = WB_CARRY_CAP + n_new - ind, dst = WB_CARRY_CAP - ind;   /* dst < src: ascending copy is safe */
    for (int i = 0; i < ind; i += 32) {
        float v = (i + lane < ind) ? row[src + i + lane] : 0.0f;
        __syncwarp();
        if (i + lane < ind) row[dst + i + lane] = v;
        __syncwarp();
    }
}

#endif /* WB_DEFRAME_KERNEL_CUH */
