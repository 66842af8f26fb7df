// sad_ws.cuh — warp-specialised, double-buffered kernel for block_size <= 17 (h <= 8), every max_disparity.
//
// Arithmetic: sad_common.cuh.  Scheduling: a CTA owns NS adjacent 32-column strips of one row band and has 24 warps
// with FIXED roles (no phase alternation, one __syncthreads per 10-row batch), registers rebalanced between the roles
// with setmaxnreg:
//   warps 0..9   walkers: warp w walks row w of the batch; its 32 lanes are (strip, disparity group) pairs (h <= 4), or
//                (row, disparity group) pairs of rows 2w and 2w+1 (h >= 5, where a batch has 2h+2 = 12..18 rows):
//                horizontal running window sums -> H[buf][row][strip][group][column] in shared memory;
//   warp 10      tail walker (modes with a tail group: the last group of the chunk, one lane per (row, strip));
//   warp 11      TMA loader (cp.async.bulk.tensor.3d + mbarrier, hardware zero fill outside the image; it also
//                replicates the left pixels);
//   warps 12..23 consumers: a warp owns 2 or 3 groups of one strip for 32 columns (lanes = columns): vertical running
//                sums with the previous rows in a REGISTER ring (shared memory carries every H value once), key-min
//                argmin; the consumers of a strip meet in ONE key per pixel through a shared-memory atomicMin.  Each
//                consumer warp also finishes a twelfth of the pixel rows of the batch before: key -> d*255/D LUT -> store.
// Producers work on batch i while consumers work on batch i-1 (H is double-buffered, tiles are requested three
// batches ahead).
//
// Two things set the speed of this kernel once the arithmetic is fixed (profiles/r02_ws_*):
//   * the ring is 10 rows long, one more than the largest window (9): the slot a row is loaded into was read for the
//     last time one row earlier, so the load lands in it directly and no register is ever moved (a 9-row ring of a
//     9-row window costs two moves per group and row: 7 % of all instructions in round 1);
//   * an SM sub-partition issues one instruction per clock for ALL its warps (warp w lives on sub-partition w % 4), so
//     the batch time is the instruction count of the most loaded sub-partition.  The roles above put 3 + 3 + (2 + tail)
//     + (2 + loader/finisher) producer warps on the four sub-partitions and the consumer table below deals the groups
//     so that every sub-partition ends up with about the same number of instructions per batch.
//
// h >= 5 (block_size 11..17): the batch is as long as the ring must be (window + 1 = 12 / 14 / 16 / 18 rows), so two H
// buffers of 33-group rows no longer fit shared memory: chunks of 17 groups on one strip, two rows per walker warp.
// h = 8 (block_size 16, 17: the reference's start-up default, params.go:13-18): window sums need 17 bits, the consumers keep
// 32-bit sums (raw packed sum + high-lane sum, keys sum*512+d) while H and the ring stay 16x2-packed.
//
// MODE = how the walker lanes are spent, i.e. which disparity ranges fill the machine (h <= 4; h >= 5 has mode 1 with one strip,
// and for h <= 7 mode 2 = 2 strips x (8 groups + tail) x 2 rows per warp for max_disparity <= 32):
//   0: 1 strip  x 32 groups + tail  = chunks of 33 groups (132 disparity slots): max_disparity 65..128, 256 in two chunks
//   1: 2 strips x 16 groups + tail  = 17 groups: max_disparity 33..64 (the reference's default range, params.go:13-18)
//   2: 3 strips x  9 groups         =  9 groups: max_disparity 17..32
//   3: 6 strips x  5 groups         =  5 groups: max_disparity <= 16
#pragma once
#include "sad_common.cuh"

#ifndef WS_REGS_SMALL
#define WS_REGS_SMALL 40    // walker registers of the h <= 4 instances (setmaxnreg)
#endif
#ifndef WS_SKIP
#define WS_SKIP 0       // developer timing experiments (results wrong): 1 = walkers idle, 2 = consumers idle, 4 = tail walker idle
#endif

namespace sadgpu {

// Consumer warp k (= warp - 12, sub-partition k % 4) -> (strip, first group, number of groups).
struct WsShare { int strip, first, ng; };
__host__ __device__ constexpr WsShare ws_share(int half, int mode, int k)
{
    if (half >= 5 && mode == 2) {       // 2 x 9 groups (max_disparity <= 32 at block_size 11..15): six warps per strip, 2 2 2 1 1 1
        const int s = k / 6, i = k % 6;
        return WsShare{s, i < 3 ? 2 * i : 3 + i, i < 3 ? 2 : 1};
    }
    if (half >= 5) {                    // 17 groups, one strip, 1 or 2 groups per warp: the sub-partitions that carry three walker /
        // tail warps (or the loader) take fewer groups
        const int t5[12] = {2, 2, 2, 2, 1, 1, 1, 2, 1, 1, 1, 1}, t6[12] = {2, 2, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1};
        const int t7[12] = {1, 2, 2, 2, 1, 2, 2, 1, 1, 1, 1, 1}, t8[12] = {2, 2, 2, 2, 1, 1, 2, 1, 1, 1, 1, 1};
        const int* t = half == 5 ? t5 : half == 6 ? t6 : half == 7 ? t7 : t8;
        int first = 0;
        for (int i = 0; i < k; ++i) first += t[i];
        return WsShare{0, first, t[k]};
    }
    if (mode == 0) {                    // 33 groups: sub-partitions 0..2 (three walker / tail warps each) take 8, sub-partition 3 takes 9
        const int first = k <= 8 ? 3 * k : 24 + 2 * (k - 8);               // 3 3 3 3  3 3 3 3  2 2 2 3
        return WsShare{0, first, (k >= 8 && k <= 10) ? 2 : 3};
    }
    if (mode == 1) {                    // 2 x 17 groups: six warps per strip, 3 3 3 3 3 2
        const int s = k / 6, i = k % 6;
        return WsShare{s, 3 * i, i == 5 ? 2 : 3};
    }
    if (mode == 2) {                    // 3 x 9 groups: four warps per strip, the 3-group warp rotates over the sub-partitions
        const int s = k / 4, i = k % 4;
        return WsShare{s, 2 * i + (i > s ? 1 : 0), i == s ? 3 : 2};
    }
    // 6 x 5 groups: two warps per strip, 3 + 2 or 2 + 3
    const int s = k / 2, i = k % 2;
    const int n0 = (s == 0 || s == 3 || s == 4) ? 3 : 2;
    return WsShare{s, i == 0 ? 0 : n0, i == 0 ? n0 : 5 - n0};
}

template <int HALF, int MODE> struct WsCfg {
    static_assert(HALF >= 0 && HALF <= 8 && MODE >= 0 && MODE <= 3 && (HALF <= 4 || MODE == 1 || (MODE == 2 && HALF <= 7)),
                  "warp-specialised kernel: block_size <= 17");
    static constexpr int WIN = 2 * HALF + 1;
    static constexpr bool WIDE = WIN * WIN * 255 >= 65536;                             // h = 8: window sums need 17 bits
    // how a candidate the reference never evaluates loses: h <= 5 a bias of 0x8000 in its running sum (sum + bias < 2^16); h = 6, 7
    // per-candidate key constants (multiplier / mask 0 and an all-ones addend); h = 8 a bias of 2^22 in its 32-bit sum
    static constexpr bool KEYC = !WIDE && WIN * WIN * 255 + 32768 >= 65536;
    static constexpr int NRW = HALF <= 4 ? 1 : 2;                                      // rows per walker warp
    static constexpr int NS = HALF >= 5 ? (MODE == 2 ? 2 : 1) : MODE == 0 ? 1 : MODE == 1 ? 2 : MODE == 2 ? 3 : 6;   // strips per CTA
    static constexpr int NGL = HALF >= 5 ? (MODE == 2 ? 8 : 16) : MODE == 0 ? 32 : MODE == 1 ? 16 : MODE == 2 ? 9 : 5;  // groups walked by the row warps
    static constexpr bool TAIL = HALF >= 5 || MODE <= 1;                               // one more group, walked by the tail warp
    static constexpr int NGC = NGL + (TAIL ? 1 : 0);                                   // groups per chunk
    static constexpr int TW = 32, TWP = 33, CW = NS * TW;                              // strip / padded strip / CTA width
    static constexpr int NSTEP = TW + 2 * HALF;                                        // steps of one strip walk
    // replicated left pixels per tile row; with two rows per walker warp (and one lane per row in the tail warp) the row stride is
    // padded to 4 mod 32 words so that the 16-byte loads of different rows fall into different banks
    static constexpr int LW0 = (CW + 2 * HALF + 3) & ~3;
    static constexpr int LW = NRW == 1 ? LW0 : LW0 + ((4 - LW0 % 32 + 32) % 32);
    static constexpr int RB = HALF <= 4 ? 10 : WIN + 1;    // rows per batch = length of the register ring (> the window): 10, or 12 / 14 / 16 / 18
    static constexpr int NWW = RB / NRW;               // row-walker warps
    static constexpr int K = 12;                       // consumer warps
    // consumers load the next row one row early: measured +7 % in the 2 x 9-group layout of block_size 11..15 (34.3 -> 31.9 us at
    // B = 15, D = 32), -1..-4 % everywhere else (headline 70.3 -> 71.3 us; the 17-group instances spill)
    static constexpr bool PFROW = HALF >= 5 && MODE == 2;
    static constexpr int NSLOT = NS * NGC;             // group slots of an H row
    // warp roles (6 warpgroups of 4 warps): producer class = warps 0..11, consumer class = warps 12..23
    static constexpr int W_AUX = NWW;                  // tail walker (the warp idles in the modes without a tail group)
    static constexpr int W_LOAD = 11;                  // tile loader
    static constexpr int W_CONS = 12;
    // finishing (key -> LUT -> store) of the batch before the one consumed.  One strip: the loader warp does it (its sub-partition
    // has the fewest instructions; measured 67 vs 73 us at D = 128).  Several strips: 20..60 rows per batch are too long a chain
    // for one warp, so item (strip, row) t belongs to consumer warp t % K (measured 46.6 -> 40.8 us at D = 64, 37.7 -> 19.3 at D = 16).
    static constexpr bool FIN_CONS = NS > 1;
    static constexpr int W_FIN = (!FIN_CONS && NWW + (TAIL ? 1 : 0) <= 10) ? 10 : W_LOAD;   // one-strip modes: a free producer warp finishes, else the loader
    static constexpr int NFI = (NS * RB + K - 1) / K;
    static constexpr int NTILE = 4;                    // tile buffers: tiles are requested three batches ahead, completed two ahead
    static constexpr int NT = 768;
    static constexpr int REGS_LAUNCH = 80, REGS_PROD = 56, REGS_CONS = 104;
    // with TMA the walkers carry no prefetch state; setmaxnreg moves registers inside the CTA's launch allocation.  A walk keeps the
    // 2h+1 old terms of both packed sums in registers: 40 registers hold a 9-wide window, the 11..17-wide ones need 48 / 56
    static constexpr int REGS_PROD_TMA = HALF <= 4 ? WS_REGS_SMALL : HALF <= 6 ? 48 : 56, REGS_CONS_TMA = 160 - REGS_PROD_TMA;
    static constexpr int OFF = walk_off(HALF);
    static constexpr int NWALKW = walk_words(HALF, NSTEP);
    static constexpr int RW = NGC - 1 + (TW / 4) * (NS - 1) + NWALKW;       // aligned right words per tile row
    static constexpr int H_BYTES = ((RB * NSLOT * TWP * 8 + 15) / 16) * 16;  // one buffer
    static constexpr int L_BYTES = RB * LW * 4;
    // TMA needs the innermost start coordinate on a 16-byte boundary: the right tile starts up to 12 bytes early (RWT words
    // per row), the raw left tile LSH bytes early (x0 is a multiple of 32, so LSH only depends on h).
    static constexpr int RWT0 = ((RW * 4 + 12 + 15) / 16) * 4;
    static constexpr int RWT = NRW == 1 ? RWT0 : RWT0 + ((12 - RWT0 % 32 + 32) % 32);      // two rows per warp: row stride 12 mod 32 words (few shared banks)
    static constexpr int R_BYTES = ((RB * RWT * 4 + 127) / 128) * 128;        // 128-byte multiple: each buffer is a TMA destination
    static constexpr int LSH = (16 - HALF % 16) % 16;
    static constexpr int LBOX = ((LSH + LW + 15) / 16) * 16;                  // TMA box width of the raw left tile (bytes)
    static constexpr int LRAW_BYTES = ((RB * LBOX + 127) / 128) * 128;
    static constexpr int PK_BYTES = RB * NS * TW * 4;                         // ONE key per pixel and row
    static constexpr int OFF_R = ((2 * H_BYTES + 127) / 128) * 128;            // TMA destinations first (128-byte aligned)
    static constexpr int OFF_LRAW = OFF_R + NTILE * R_BYTES;
    static constexpr int OFF_L = OFF_LRAW + NTILE * LRAW_BYTES;
    static constexpr int OFF_PK = OFF_L + NTILE * L_BYTES;
    static constexpr int OFF_LUT = OFF_PK + 2 * PK_BYTES;
    static constexpr int OFF_MBAR = OFF_LUT + 1040;
    static constexpr int SMEM = OFF_MBAR + 64;
    static_assert(WIN < RB, "the ring must be longer than the window");
    static_assert(NT * REGS_LAUNCH <= 65536 && 384 * REGS_PROD + 384 * REGS_CONS <= NT * REGS_LAUNCH &&
                  384 * REGS_PROD_TMA + 384 * REGS_CONS_TMA <= NT * REGS_LAUNCH, "register budget");
    static_assert(NRW * NS * NGL <= 32 && (!TAIL || RB * NS <= 32) && RB % NRW == 0 && NWW + (TAIL ? 1 : 0) <= W_LOAD, "walker lanes and warps");
    static_assert(RWT * 4 <= 256 && LBOX <= 256 && RB <= 256, "TMA box");
    static_assert(SMEM <= 232448, "shared memory");
};

// Consumer warp: NG groups (first group gf of strip s) x 32 columns; one barrier per batch.
template <int HALF, int MODE, int NG>
__device__ __forceinline__ void ws_consume(const FastArgs& a, const uint2* __restrict__ Hs, uint32_t* __restrict__ pk,
                                           const uint8_t* __restrict__ lut, int kB, int s, int gf, int lane, int frame,
                                           int x0, int g0, int r0, int yb1, int nb)
{
    using C = WsCfg<HALF, MODE>;
    constexpr int WIN = C::WIN, TW = C::TW, TWP = C::TWP, RB = C::RB, NS = C::NS;
    constexpr int HBUF = C::H_BYTES / 8, HROW = C::NSLOT * TWP, PKBUF = RB * C::NS * TW;
    const int xB = x0 + s * TW + lane;
    // finishing share of this warp: items t = kB, kB + K, ... (strip t / RB, row t % RB) of the batch before the one consumed
    uint8_t* __restrict__ Og = a.out + (long long)frame * a.frameOut;
    int fs[C::NFI], frb[C::NFI];
    bool fx[C::NFI], fz[C::NFI];
#pragma unroll
    for (int i = 0; i < (C::FIN_CONS ? C::NFI : 0); ++i) {
        const int t = kB + C::K * i;
        fs[i] = t / RB; frb[i] = t - fs[i] * RB;
        const int x = x0 + fs[i] * TW + lane;
        fx[i] = t < NS * RB && x < a.W;
        fz[i] = x < HALF;                                      // X < h: both windows clamp, d = 0 wins (sad.go:212-218)
    }
    auto finish = [&](int batch) {
        uint32_t* pkb = pk + (batch & 1) * PKBUF + lane;
        const int row0 = r0 + batch * RB - HALF;               // image row of item row 0
        const int rb_lo = 2 * HALF - batch * RB, rb_hi = yb1 - row0;   // output rows of this batch: rb in [rb_lo, rb_hi)
#pragma unroll
        for (int i = 0; i < C::NFI; ++i) {
            if (kB + C::K * i >= NS * RB) continue;            // warp-uniform
            uint32_t* q = pkb + (frb[i] * NS + fs[i]) * TW;
            uint32_t best = *q;
            *q = 0xFFFFFFFFu;                                  // re-arm for the batch after next
            if (fz[i]) best = 0;
            const bool ok = fx[i] && frb[i] >= rb_lo && frb[i] < rb_hi;
            const int x = x0 + fs[i] * TW + lane, y = row0 + frb[i];
            if (a.NC == 1) { const uint8_t v = lut[min(best & (C::WIDE ? 511u : 0xFFFFu), 1039u)]; if (ok) Og[(long long)y * a.pitchOut + x] = v; }
            else if (ok) atomicMin(a.gkey + ((long long)frame * a.H + y) * a.W + x, C::WIDE ? best : (((best >> 16) << 9) | (best & 511u)));
        }
    };
    // Narrow (h <= 7): VE / VO are 16x2-packed window sums.  Wide (h = 8): VE / VO are the RAW packed sums (low lane + 65536 * high
    // lane, mod 2^32) and V3 / V2 the high-lane sums alone; keys are sum * 512 + d, the low lane's being VE*512 - V3*2^25 + d.
    uint32_t VE[NG], VO[NG], V3[NG], V2[NG], ringE[RB][NG], ringO[RB][NG];
    uint32_t mE[NG], aE[NG], nE[NG], oE[NG], mO[NG], aO[NG], nO[NG], oO[NG];            // h = 6, 7: per-candidate key constants
    const uint32_t kc16 = a.k65536, kchi = a.k65536 * 0xFFFFu;
#pragma unroll
    for (int j = 0; j < NG; ++j) {
        const int dbase = 4 * (g0 + gf + j);
        const int dmax = min(a.D, xB - HALF);                 // largest evaluated disparity of this column (sad.go:64-67, :212-218)
        if (C::WIDE) {
            // bias: a never-evaluated candidate starts 2^22 above every real sum (245 055): its key carries 2^31 (the high lane's
            // bias drops out of the low lane's key: 2^22 * 2^25 = 0 mod 2^32)
            VE[j] = dbase + 3 > dmax ? 1u << 22 : 0u; V3[j] = dbase + 1 > dmax ? 1u << 22 : 0u;
            VO[j] = dbase + 2 > dmax ? 1u << 22 : 0u; V2[j] = dbase + 0 > dmax ? 1u << 22 : 0u;
        } else if (C::KEYC) {
            const bool v3 = dbase + 3 <= dmax, v1 = dbase + 1 <= dmax, v2 = dbase + 2 <= dmax, v0 = dbase <= dmax;
            mE[j] = v3 ? kc16 : 0u; aE[j] = v3 ? (uint32_t)(dbase + 3) : 0xFFFFFFFFu;
            nE[j] = v1 ? kchi : 0u; oE[j] = v1 ? (uint32_t)(dbase + 1) : 0xFFFFFFFFu;
            mO[j] = v2 ? kc16 : 0u; aO[j] = v2 ? (uint32_t)(dbase + 2) : 0xFFFFFFFFu;
            nO[j] = v0 ? kchi : 0u; oO[j] = v0 ? (uint32_t)dbase : 0xFFFFFFFFu;
            VE[j] = 0; VO[j] = 0; V3[j] = 0; V2[j] = 0;
        } else {
            const uint32_t iE = (dbase + 3 > dmax ? 0x0000FFFFu : 0u) | (dbase + 1 > dmax ? 0xFFFF0000u : 0u);
            const uint32_t iO = (dbase + 2 > dmax ? 0x0000FFFFu : 0u) | (dbase + 0 > dmax ? 0xFFFF0000u : 0u);
            VE[j] = iE & 0x80008000u;                         // bias: never-evaluated candidates lose
            VO[j] = iO & 0x80008000u;
            V3[j] = 0; V2[j] = 0;
        }
#pragma unroll
        for (int r = 0; r < RB; ++r) { ringE[r][j] = 0; ringO[r][j] = 0; }
    }
    const uint32_t keybase = 4u * (uint32_t)(g0 + gf);
    // key constants held in registers (opaque to the compiler, so that the keys stay IMADs / LOP3s with an immediate index); only
    // the pair the instance uses is materialised
    const uint32_t k16 = (C::WIDE || C::KEYC) ? 0u : opaque(a.k65536), mhi = (C::WIDE || C::KEYC) ? 0u : opaque(a.k65536 * 0xFFFFu);
    const uint32_t k512 = C::WIDE ? opaque(a.k65536 >> 7) : 0u, m25 = C::WIDE ? opaque(0u - (a.k65536 << 9)) : 0u;     // wide keys: 512 and -2^25
    const uint2* Hbase = Hs + (s * C::NGC + gf) * TWP + lane;
    uint32_t* pkbase = pk + s * TW + lane;
    for (int it = 0; it < nb + 2; ++it) {
        if (C::FIN_CONS && it >= 2) finish(it - 2);
        if (it >= 1 && it <= nb && !(WS_SKIP & 2)) {
            const int batch = it - 1;
            const uint2* Hp = Hbase + (batch & 1) * HBUF;
            uint32_t* pkb = pkbase + (batch & 1) * PKBUF;
            // PFROW: the loads of row rb + 1 are issued before the atomicMin of row rb (a shared-memory load cannot be hoisted over it)
            uint2 nn[NG];
#pragma unroll
            for (int j = 0; j < NG; ++j) nn[j] = C::PFROW ? Hp[j * TWP] : make_uint2(0u, 0u);
#pragma unroll
            for (int rb = 0; rb < RB; ++rb) {
                uint32_t best = 0xFFFFFFFFu;
                uint2 nc[NG];
#pragma unroll
                for (int j = 0; j < NG; ++j) {
                    nc[j] = C::PFROW ? nn[j] : Hp[rb * HROW + j * TWP];
                    if (C::PFROW && rb + 1 < RB) nn[j] = Hp[(rb + 1) * HROW + j * TWP];
                }
#pragma unroll
                for (int j = 0; j < NG; ++j) {
                    // the slot of row rb was read for the last time one row ago (WIN < RB): the load lands in it directly
                    const uint2 n = nc[j];
                    uint32_t kEl, kEh, kOl, kOh;
                    if (C::WIDE) {
                        // packed difference with the low lane biased by 0x8000: it never borrows from the high lane, so the
                        // arithmetic shift yields the signed high-lane difference
                        const uint32_t dE = n.x - ringE[(rb + RB - WIN) % RB][j] + 0x8000u;
                        const uint32_t dO = n.y - ringO[(rb + RB - WIN) % RB][j] + 0x8000u;
                        VE[j] = VE[j] + dE - 0x8000u; V3[j] += (uint32_t)((int)dE >> 16);
                        VO[j] = VO[j] + dO - 0x8000u; V2[j] += (uint32_t)((int)dO >> 16);
                        kEl = VE[j] * k512 + (V3[j] * m25 + (4u * j + 3u)); kEh = V3[j] * k512 + (4u * j + 1u);
                        kOl = VO[j] * k512 + (V2[j] * m25 + (4u * j + 2u)); kOh = V2[j] * k512 + (4u * j + 0u);
                    } else if (C::KEYC) {
                        VE[j] = VE[j] + n.x - ringE[(rb + RB - WIN) % RB][j];
                        VO[j] = VO[j] + n.y - ringO[(rb + RB - WIN) % RB][j];
                        kEl = VE[j] * mE[j] + aE[j]; kEh = (VE[j] & nE[j]) | oE[j];           // absolute disparities in the addends
                        kOl = VO[j] * mO[j] + aO[j]; kOh = (VO[j] & nO[j]) | oO[j];
                    } else {
                        VE[j] = VE[j] + n.x - ringE[(rb + RB - WIN) % RB][j];
                        VO[j] = VO[j] + n.y - ringO[(rb + RB - WIN) % RB][j];
                        kEl = key_lo(VE[j], k16, 4u * j + 3u);
                        kEh = key_hi(VE[j], mhi, 4u * j + 1u);
                        kOl = key_lo(VO[j], k16, 4u * j + 2u);
                        kOh = key_hi(VO[j], mhi, 4u * j + 0u);
                    }
                    ringE[rb][j] = n.x; ringO[rb][j] = n.y;
                    best = min(best, min(kEl, kEh));
                    best = min(best, min(kOl, kOh));
                }
                atomicMin(pkb + rb * C::NS * TW, C::KEYC ? best : best + keybase);    // rows that are not output rows are filtered by the finisher
            }
        }
        __syncthreads();
    }
}

template <int HALF, int MODE, bool TMA>
__global__ void __launch_bounds__(WsCfg<HALF, MODE>::NT, 1) sad_ws_kernel(const __grid_constant__ FastArgs a)
{
    using C = WsCfg<HALF, MODE>;
    constexpr int TW = C::TW, TWP = C::TWP, RB = C::RB, NGC = C::NGC, NS = C::NS, NGL = C::NGL;
    extern __shared__ __align__(128) unsigned char smem[];
    uint2* Hs = reinterpret_cast<uint2*>(smem);                                   // [2][RB][NS][NGC][TWP]
    uint32_t* Lrep = reinterpret_cast<uint32_t*>(smem + C::OFF_L);               // [NTILE][RB][LW]
    uint32_t* Ral = reinterpret_cast<uint32_t*>(smem + C::OFF_R);                // [NTILE][RB][RWT]
    uint32_t* pk = reinterpret_cast<uint32_t*>(smem + C::OFF_PK);                // [2][RB][NS][TW]
    uint8_t* lut = smem + C::OFF_LUT;
    constexpr int HBUF = C::H_BYTES / 8, HROW = C::NSLOT * TWP, LBUF = RB * C::LW, RBUF = C::R_BYTES / 4, PKBUF = RB * NS * TW;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int frame = blockIdx.z / a.NC, chunk = blockIdx.z - frame * a.NC;
    const int x0 = blockIdx.x * C::CW;
    const int yb0 = a.y0 + blockIdx.y * a.BH;
    const int yb1 = min(a.y1, yb0 + a.BH);
    const int g0 = chunk * NGC;
    if (yb0 >= yb1) return;
    // a chunk none of whose disparities is a candidate anywhere in this strip (d > X-h for every column, sad.go:64-67 + :212-218)
    // has nothing to contribute: chunk 0 always runs and writes every pixel
    if (g0 > 0 && min(x0 + C::CW, a.W) - 1 - HALF < 4 * g0) return;
    const int r0 = yb0 - HALF;
    const int nb = ((yb1 - yb0) + 2 * HALF + RB - 1) / RB;
    // image row r0 + batch*RB + rb is row rb of `batch`; the walk of strip s starts at column x0 + 32 s - h
    const int xr0 = x0 - HALF - 3 - 4 * (g0 + NGC - 1) - C::OFF;              // image column of right-tile word 0 (multiple of 4)
    const int xr0a = xr0 - (((xr0 % 16) + 16) % 16);                          // tile rows start 16-byte aligned (TMA rule), same layout without TMA
    const int rext = (xr0 - xr0a) >> 2;                                       // words to skip at the start of a tile row

    for (int d = tid; d < 1040; d += C::NT) lut[d] = d <= a.D ? (uint8_t)((d * 255) / a.D) : 0;
    for (int idx = tid; idx < 2 * PKBUF; idx += C::NT) pk[idx] = 0xFFFFFFFFu;

    if (warp < C::W_CONS) {
        // ======================= producer warpgroups (warps 0..11) =======================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" :: "n"(TMA ? C::REGS_PROD_TMA : C::REGS_PROD));
        // Finishing (one-strip modes; batch it-2): one key per pixel -> LUT -> store, and the key is re-armed.  All key loads first,
        // then the LUT loads, then the stores: the warp that finishes is a single chain per batch, its latency must stay short.
        auto finish = [&](int batch) {
            uint8_t* __restrict__ Og = a.out + (long long)frame * a.frameOut;
            const bool xin = x0 + lane < a.W;
            const bool zero0 = x0 == 0 && lane < HALF;                          // X < h: both windows clamp, d = 0 wins (sad.go:212-218)
            uint32_t* pkb = pk + (batch & 1) * PKBUF + lane;
            const int row0 = r0 + batch * RB - HALF;                           // image row of item row 0
            const int rb_lo = 2 * HALF - batch * RB, rb_hi = yb1 - row0;       // output rows of this batch: rb in [rb_lo, rb_hi)
            uint8_t* Orow = Og + (long long)row0 * a.pitchOut + x0 + lane;
            const long long grow = ((long long)frame * a.H + row0) * a.W + x0 + lane;       // key-map index of item row 0 (chunked ranges)
            constexpr int FG = 6;
#pragma unroll
            for (int r0g = 0; r0g < RB; r0g += FG) {
                uint32_t best[FG];
#pragma unroll
                for (int g = 0; g < FG; ++g) if (r0g + g < RB) best[g] = pkb[(r0g + g) * TW];
#pragma unroll
                for (int g = 0; g < FG; ++g) if (r0g + g < RB) pkb[(r0g + g) * TW] = 0xFFFFFFFFu;   // re-arm for the batch after next
                if (a.NC == 1) {
                    uint8_t v[FG];
#pragma unroll
                    for (int g = 0; g < FG; ++g) if (r0g + g < RB) v[g] = lut[min((zero0 ? 0u : best[g]) & (C::WIDE ? 511u : 0xFFFFu), 1039u)];
#pragma unroll
                    for (int g = 0; g < FG; ++g)
                        if (r0g + g < RB && xin && r0g + g >= rb_lo && r0g + g < rb_hi) Orow[(r0g + g) * a.pitchOut] = v[g];
                } else {
#pragma unroll
                    for (int g = 0; g < FG; ++g)
                        if (r0g + g < RB && xin && r0g + g >= rb_lo && r0g + g < rb_hi) {
                            const uint32_t b = zero0 ? 0u : best[g];
                            atomicMin(a.gkey + grow + (long long)(r0g + g) * a.W, C::WIDE ? b : (((b >> 16) << 9) | (b & 511u)));
                        }
                }
            }
        };
        if (warp != C::W_LOAD && warp >= C::NWW + (C::TAIL ? 1 : 0)) {
            // ---- producer-class warps without a walk: warp 10 finishes when it is free (h >= 5), the others keep the barrier count ----
            __syncthreads();
            for (int it = 0; it < nb + 2; ++it) {
                if (!C::FIN_CONS && C::W_FIN != C::W_LOAD && warp == C::W_FIN && it >= 2) finish(it - 2);
                __syncthreads();
            }
        } else if (warp == C::W_LOAD) {
            // ---- loader (TMA, warp 11): two cp.async.bulk.tensor per batch (raw left rows, aligned right rows), completion on
            //      an mbarrier, then the left pixels are replicated into Lrep.  Without TMA the walkers prefetch their own
            //      rows and this warp only keeps the barrier count (and finishes, in the one-strip modes). ----
            uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + C::OFF_MBAR);
            const uint32_t mbar0 = (uint32_t)__cvta_generic_to_shared(mbar);
            auto request = [&](int batch) {                 // two bulk tensor copies, completion counted on the buffer's mbarrier
                const int tb = batch % C::NTILE;
                const uint32_t bar = mbar0 + 8 * tb;
                const uint32_t dstR = (uint32_t)__cvta_generic_to_shared(smem + C::OFF_R + tb * C::R_BYTES);
                const uint32_t dstL = (uint32_t)__cvta_generic_to_shared(smem + C::OFF_LRAW + tb * C::LRAW_BYTES);
                const int y = r0 + batch * RB;
                if (lane == 0) {
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(RB * C::RWT * 4 + RB * C::LBOX) : "memory");
                    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                                 :: "r"(dstR), "l"(&a.tmapR), "r"(xr0a), "r"(y), "r"(frame), "r"(bar) : "memory");
                    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                                 :: "r"(dstL), "l"(&a.tmapL), "r"(x0 - HALF - C::LSH), "r"(y), "r"(frame), "r"(bar) : "memory");
                }
            };
            auto complete = [&](int batch) {                // wait for the copies of `batch`, then replicate its left pixels
                const int tb = batch % C::NTILE;
                const uint32_t bar = mbar0 + 8 * tb;
                // bounded wait on the phase of this buffer's (batch / NTILE)-th use; a stuck copy traps instead of hanging
                const uint32_t parity = (uint32_t)(batch / C::NTILE) & 1u;
                uint32_t done = 0;
                for (int spin = 0; spin < (1 << 24) && !done; ++spin)
                    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
                if (!done) __trap();
                // four pixels per lane and step: one (funnel-shifted) raw word -> four replicated words, one 16-byte store
                const uint32_t* raw = reinterpret_cast<const uint32_t*>(smem + C::OFF_LRAW + tb * C::LRAW_BYTES);
                uint4* Ld = reinterpret_cast<uint4*>(Lrep + tb * LBUF);
                constexpr int LQ = C::LW / 4, NIT = (RB * LQ + 31) / 32;
#pragma unroll
                for (int k0 = 0; k0 < NIT; k0 += 4) {                // four steps at a time: all loads first (one warp, one chain per batch)
                    uint32_t v[4], v2[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int idx = lane + 32 * (k0 + k);
                        v[k] = 0; v2[k] = 0;
                        if (k0 + k < NIT && idx < RB * LQ) {
                            const int rb = idx / LQ, q = idx - rb * LQ;
                            const uint32_t* p = raw + rb * (C::LBOX / 4) + (C::LSH >> 2) + q;
                            v[k] = p[0];
                            if (C::LSH & 3) v2[k] = p[1];
                        }
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int idx = lane + 32 * (k0 + k);
                        if (k0 + k < NIT && idx < RB * LQ) {
                            const uint32_t w = (C::LSH & 3) ? __funnelshift_r(v[k], v2[k], 8 * (C::LSH & 3)) : v[k];
                            Ld[idx] = make_uint4(__byte_perm(w, 0u, 0x0000), __byte_perm(w, 0u, 0x1111), __byte_perm(w, 0u, 0x2222), __byte_perm(w, 0u, 0x3333));
                        }
                    }
                }
            };
            constexpr bool loads = TMA;
            if (loads) {
                if (lane == 0) {
#pragma unroll
                    for (int t = 0; t < C::NTILE; ++t) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(mbar0 + 8 * t));
                    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
                }
                __syncwarp();
                request(0);
                if (nb > 1) request(1);
                if (nb > 2) request(2);
                complete(0);
                if (nb > 1) complete(1);
            }
            __syncthreads();
            for (int it = 0; it < nb + 2; ++it) {
                if (loads) {
                    if (it + 3 < nb) request(it + 3);              // buffer (it+3)%4 was last read in iteration it-1
                    if (it + 2 < nb) complete(it + 2);             // requested one iteration ago: already landed
                }
                if (!C::FIN_CONS && C::W_FIN == C::W_LOAD && it >= 2) finish(it - 2);
                __syncthreads();
            }
        } else {
            // ---- walkers: warp w < RB walks row w; lane = (strip s, group gl) for the NGL groups the row warps take.
            //      The tail warp walks the last group of the chunk for every (row, strip).  Without TMA a row warp also
            //      prefetches its row of the batch after next (L pixels replicated, R as aligned words). ----
            const bool tail = C::TAIL && warp == C::W_AUX;
            const int unit = C::NRW * NS == 1 ? 0 : lane / NGL;                 // row warps: lane = ((row of the warp, strip), group)
            const int ws = tail ? lane % NS : unit % NS;                        // strip of this lane
            const int rb = tail ? lane / NS : warp * C::NRW + unit / NS;        // row of the batch
            const int gl = tail ? NGC - 1 : lane - unit * NGL;
            const bool act = tail ? lane < RB * NS : lane < C::NRW * NS * NGL;
            const int nvalid = a.W - (x0 + ws * TW - HALF);                     // steps of this lane's walk inside the image
            const bool edge = a.W - (x0 + (NS - 1) * TW - HALF) < C::NSTEP;     // warp-uniform: some strip of the CTA touches x >= W
            const uint8_t* __restrict__ Lg = a.L + (long long)frame * a.frameL;
            const uint8_t* __restrict__ Rg = a.R + (long long)frame * a.frameR;
            constexpr int NLQ = (C::LW + 31) / 32, NRQ = (C::RWT + 31) / 32;
            int lx[NLQ], rx[NRQ], rmode[NRQ];            // column of each slot of this lane; -1 / mode 0 = zero
            uint32_t vl[C::NRW][NLQ], vr[C::NRW][NRQ];
            const bool self_load = !tail && !TMA;
            if (self_load) {
#pragma unroll
                for (int q = 0; q < NLQ; ++q) {
                    const int i = lane + 32 * q, x = x0 - HALF + i;
                    lx[q] = (i < C::LW && (unsigned)x < (unsigned)a.W) ? x : -1;
                }
#pragma unroll
                for (int q = 0; q < NRQ; ++q) {
                    const int j = lane + 32 * q, x = xr0a + 4 * j;
                    const bool in = j < C::RWT && x + 3 >= 0 && x < a.W;
                    rx[q] = x;
                    rmode[q] = !in ? 0 : (a.aligned && x >= 0 && x + 3 < a.W) ? 1 : 2;
                }
            }
            auto issue = [&](int batch) {               // global loads of the NRW rows of this warp in `batch` (warp-uniform row tests)
#pragma unroll
                for (int rr = 0; rr < C::NRW; ++rr) {
                    const int y = r0 + batch * RB + warp * C::NRW + rr;
                    const bool yin = (unsigned)y < (unsigned)a.H;
                    const uint8_t* pl = Lg + (size_t)(yin ? y : 0) * a.pitchL;
                    const uint8_t* pr = Rg + (size_t)(yin ? y : 0) * a.pitchR;
#pragma unroll
                    for (int q = 0; q < NLQ; ++q) { vl[rr][q] = 0; if (yin && lx[q] >= 0) vl[rr][q] = pl[lx[q]]; }
#pragma unroll
                    for (int q = 0; q < NRQ; ++q) {
                        uint32_t v = 0;
                        if (yin && rmode[q] == 1) v = *reinterpret_cast<const uint32_t*>(pr + rx[q]);
                        else if (yin && rmode[q] == 2) {
#pragma unroll
                            for (int b = 0; b < 4; ++b)
                                if ((unsigned)(rx[q] + b) < (unsigned)a.W) v |= (uint32_t)pr[rx[q] + b] << (8 * b);
                        }
                        vr[rr][q] = v;
                    }
                }
            };
            auto commit = [&](int batch) {
#pragma unroll
                for (int rr = 0; rr < C::NRW; ++rr) {
                    uint32_t* Ld = Lrep + (batch % C::NTILE) * LBUF + (warp * C::NRW + rr) * C::LW;
                    uint32_t* Rd = Ral + (batch % C::NTILE) * RBUF + (warp * C::NRW + rr) * C::RWT;
#pragma unroll
                    for (int q = 0; q < NLQ; ++q) { const int i = lane + 32 * q; if (i < C::LW) Ld[i] = vl[rr][q] * 0x01010101u; }
#pragma unroll
                    for (int q = 0; q < NRQ; ++q) { const int j = lane + 32 * q; if (j < C::RWT) Rd[j] = vr[rr][q]; }
                }
            };
            if (self_load) {
                issue(0); commit(0);
                if (nb > 1) { issue(1); commit(1); }
            }
            __syncthreads();
            for (int it = 0; it < nb + 2; ++it) {
                const bool pre = self_load && it + 2 < nb;
                if (pre) issue(it + 2);
                if (it < nb && act && !(WS_SKIP & 1) && !((WS_SKIP & 4) && tail)) {
                    const int buf = it & 1, tb = it % C::NTILE;
                    const uint32_t* Lr = Lrep + tb * LBUF + rb * C::LW + ws * TW;
                    const uint32_t* Rr = Ral + tb * RBUF + rb * C::RWT + rext + (NGC - 1 - gl) + ws * (TW / 4);
                    uint2* Hout = Hs + buf * HBUF + rb * HROW + (ws * NGC + gl) * TWP;
                    if (!edge) sad_walk<HALF, TW, false>(Lr, Rr, Hout, nvalid);
                    else       sad_walk<HALF, TW, true>(Lr, Rr, Hout, nvalid);
                }
                if (pre) commit(it + 2);               // tile buffer (it+2)%4 was last read in iteration it-2
                __syncthreads();
            }
        }
    } else {
        // ======================= consumer warpgroups (warps 12..23) =======================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" :: "n"(TMA ? C::REGS_CONS_TMA : C::REGS_CONS));
        __syncthreads();
        // ---- consumers: vertical running sums (register ring) + argmin keys for 2 or 3 groups x 32 columns ----
        const int kB = warp - C::W_CONS;
        WsShare sh = ws_share(HALF, MODE, 0);
#pragma unroll
        for (int k = 1; k < C::K; ++k)
            if (k == kB) sh = ws_share(HALF, MODE, k);
        if (HALF <= 4) {                                               // 2 or 3 groups per warp
            if (sh.ng == 3) ws_consume<HALF, MODE, 3>(a, Hs, pk, lut, kB, sh.strip, sh.first, lane, frame, x0, g0, r0, yb1, nb);
            else            ws_consume<HALF, MODE, 2>(a, Hs, pk, lut, kB, sh.strip, sh.first, lane, frame, x0, g0, r0, yb1, nb);
        } else {                                                       // 1 or 2 groups per warp (the ring is 12..18 rows long)
            if (sh.ng == 2) ws_consume<HALF, MODE, 2>(a, Hs, pk, lut, kB, sh.strip, sh.first, lane, frame, x0, g0, r0, yb1, nb);
            else            ws_consume<HALF, MODE, 1>(a, Hs, pk, lut, kB, sh.strip, sh.first, lane, frame, x0, g0, r0, yb1, nb);
        }
    }
}

}  // namespace sadgpu
