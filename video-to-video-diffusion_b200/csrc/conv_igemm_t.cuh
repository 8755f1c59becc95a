// Operand-swapped implicit-GEMM convolution for Cout == 128 layers (U-Net level 0, VAE 128-channel stages).
//
// With M = positions(128) x N = Cout(128) the tcgen05 SS-mode MMA reads 4 KB of A and 4 KB of B from shared
// memory every 64 cycles -- the shared-memory pipe, not the tensor pipe, is the limit (measured: ~0.9 PFLOP/s
// against ~1.45 for the N = 256 layers).  Here the roles are swapped: the weight tile (128 channels) is the
// UMMA "A" operand and TWO position boxes (256 rows) form the "B" operand, i.e. D^T[channel][position],
// M = 128 x N = 256 x K = 16: 12 KB of operands per 128 cycles, the same ratio as the N = 256 kernel.
// The accumulator comes out transposed (TMEM lane = channel, column = position), so the epilogue
//   * adds the per-lane bias, sums each channel over the valid positions in registers (GroupNorm statistics need
//     only log2(cpg) shuffles because a group is cpg adjacent lanes),
//   * transposes 32x32 blocks through a per-warp shared-memory tile and writes NDHWC fp16 rows with 16-byte stores.
// Everything else (TMA halo boxes, tap tables, parity classes, barriers, persistent scheduling, TMEM double
// buffering) is the scheme of conv_igemm.cuh; ConvParams is shared.
#pragma once
#include "conv_igemm.cuh"

namespace b2v {

struct ConvCfgT {
  static constexpr int W_BYTES = 128 * 128;      // weight tile: 128 channels x 64 k
  static constexpr int P_BYTES = 2 * 128 * 128;  // two position boxes x 64 k
  static constexpr int STAGE = W_BYTES + P_BYTES;
  static constexpr int NSTAGE = 4;
  static constexpr int TM_COLS = 512;
  static constexpr int XPOSE = 4 * 32 * 33 * 4;  // transpose tiles: 8 warps x 32x32 fp16 (MODE 0) or 4 warps x 32x33 fp32 (MODE 1)
  static constexpr int STATS = 2048;             // [2 epilogue warp groups][<=128 groups][2] floats
  static constexpr int SMEM = NSTAGE * STAGE + 1024 + 256 + STATS + XPOSE;
  // epilogue warps: MODE 0 runs two per TMEM lane quarter (each takes four of the tile's eight 32-position chunks)
  static constexpr int ew(int mode) { return mode == 0 ? 8 : 4; }
  static constexpr int threads(int mode) { return 64 + 32 * ew(mode); }
};

// MODE 0: convolution (above).  MODE 1 ("tap GEMM", narrow heads Cout <= 16): the 128 "channel" rows are
// (kh, kw, cout) triples of a 3x3x3 filter, the positions are 128-wide runs of one depth slice with no in-plane halo,
// the three depth taps are accumulated in the K loop (depth-shifted TMA loads, zero-filled outside the volume), and
// the raw fp32 products P[(kh*3+kw)*Cout+co][position] are stored row-wise; head_stencil_kernel then sums the 9
// in-plane shifted rows.  The input is read 3 times (L2 hits) instead of 27 (the N=16 implicit GEMM was L2-bound at
// ~7 TB/s) and P is 9*Cout rows (a 27*Cout-row P with the input read once spilled the L2: 382 MB for the U-Net head).
template <int MODE>
__global__ void __launch_bounds__(ConvCfgT::threads(MODE), 1) conv_igemm_t_kernel(const __grid_constant__ ConvParams p) {
  using Cfg = ConvCfgT;
  constexpr int NSTAGE = Cfg::NSTAGE;
  constexpr int EW = Cfg::ew(MODE);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + NSTAGE * Cfg::STAGE);
  uint64_t* empty = full + NSTAGE;
  uint64_t* tfull = empty + NSTAGE;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  float* sstat = reinterpret_cast<float*>(smem + NSTAGE * Cfg::STAGE + 256);
  __half* xpose = reinterpret_cast<__half*>(smem + NSTAGE * Cfg::STAGE + 256 + Cfg::STATS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_tiles = p.tiles_w * p.tiles_h * p.tiles_d * p.batch;  // even per sample (checked on the host)
  const int pairs = (m_tiles + 1) >> 1;
  const int total = pairs * (MODE ? p.n_tiles : p.nclass);
  const int chunks = p.src_chunks0 + p.src_chunks1;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&p.tmA[0]);
    tma_prefetch_desc(&p.tmB);
    for (int i = 0; i < NSTAGE; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], EW);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::TM_COLS);
  for (int i = threadIdx.x; i < 512; i += blockDim.x) sstat[i] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // prologue done (barriers, TMEM, descriptors): let the next kernel start its own, then wait for our producer
  pdl_trigger();
  pdl_wait();

  if (warp == 0) {
    if (lane == 0) {
      const uint32_t a_bytes = (uint32_t)p.rows_valid * 128u;
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const int cls = MODE ? 0 : tile / pairs;
        const int n0 = MODE ? (tile / pairs) * 128 : 0;
        const int pm = tile % pairs;
        int w0[2], h0[2], d0[2], nb[2];
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          int m = 2 * pm + hh;
          w0[hh] = (m % p.tiles_w) * p.bw;
          m /= p.tiles_w;
          h0[hh] = (m % p.tiles_h) * p.bh;
          m /= p.tiles_h;
          d0[hh] = (m % p.tiles_d) * p.bd;
          nb[hh] = m / p.tiles_d;
        }
        for (int t = 0; t < p.ntaps; ++t) {
          const int tg = cls * p.ntaps + t;
          const int32_t tp = p.taps[tg];
          const int map = tp >> 24;
          const int od = ((tp >> 16) & 0xff) - 8, oh = ((tp >> 8) & 0xff) - 8, ow = (tp & 0xff) - 8;
          for (int c = 0; c < chunks; ++c) {
            const int src = (c >= p.src_chunks0) ? 1 : 0;
            const int cc = src ? (c - p.src_chunks0) : c;
            mbar_wait(&empty[stage], phase ^ 1);
            mbar_arrive_expect_tx(&full[stage], 2u * a_bytes + (uint32_t)Cfg::W_BYTES);
            uint8_t* sa = smem + stage * Cfg::STAGE;
            tma_load_3d(sa, &p.tmB, &full[stage], c * 64, n0, tg);
            tma_load_5d(sa + Cfg::W_BYTES, &p.tmA[map + src], &full[stage], cc * 64, w0[0] + ow, h0[0] + oh,
                        d0[0] + od, nb[0]);
            tma_load_5d(sa + Cfg::W_BYTES + 16384, &p.tmA[map + src], &full[stage], cc * 64, w0[1] + ow, h0[1] + oh,
                        d0[1] + od, nb[1]);
            if (++stage == NSTAGE) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(128, 256, 0);
      const int ksteps = p.ntaps * chunks;
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * 256);
        for (int k = 0; k < ksteps; ++k) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE);
          const uint64_t wdesc = umma_desc_sw128(sa);                 // 128 channel rows
          const uint64_t pdesc = umma_desc_sw128(sa + Cfg::W_BYTES);  // 256 position rows
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_f16(tmem_d, wdesc + (uint64_t)(kk * 2), pdesc + (uint64_t)(kk * 2), idesc, (k | kk) ? 1u : 0u);
          umma_commit(&empty[stage]);
          if (++stage == NSTAGE) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tfull[acc]);
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue: lane = output channel
    const int q = warp & 3;
    const int ch = q * 32 + lane;
    const int et = threadIdx.x - 64;
    const int eg = (warp - 2) >> 2;      // 0 / 1: which four chunks of the tile this warp handles (MODE 0)
    float* gstat = sstat + eg * 256;     // this warp group's table (a group's entry has ONE writer lane: fixed order)
    constexpr int CHUNKS = 32 / EW;      // chunks per warp: 4 with eight epilogue warps, 8 with four
    const int ng = 128 / p.cpg;
    const float bias_c = MODE ? 0.f : __ldg(p.bias + ch);
    __half* xp = xpose + (warp - 2) * 1024;  // 32 positions x 32 channels (MODE 0)
    __half* outp = reinterpret_cast<__half*>(p.out);
    int cur_nb = -1;
    auto flush = [&]() {
      asm volatile("bar.sync 1, %0;" ::"n"(32 * EW) : "memory");
      if (et < 2 * ng) {
        const float val = sstat[et] + sstat[256 + et];  // the two warp groups' tables, in index order
        sstat[et] = sstat[256 + et] = 0.f;
        stat_add(p.stats + ((size_t)cur_nb * p.groups + (et >> 1)) * 2 + (et & 1), val);
      }
      asm volatile("bar.sync 1, %0;" ::"n"(32 * EW) : "memory");
    };
    int it = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int cls = MODE ? 0 : tile / pairs;
      const int pm = tile % pairs;
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 256);
      if constexpr (MODE == 1) {
        // raw rows: TMEM lane = (kh, kw, cout) row, columns = positions.  A 32x32 fp32 block is transposed through
        // shared memory (row pitch 33: conflict-free both ways) so that every store instruction writes one full
        // 128-byte line of ONE row (16-byte pieces scattered over many rows made DRAM writes crawl once P outgrew L2)
        // logical row r of a 128-row tile sits in TMEM lane (r % 4) * 32 + r / 4 (weights are packed that way), so
        // the valid rows of a narrow head (9 for Cout = 1) are spread over all four epilogue warps
        const int tbase = (tile / pairs) * 128;
        const int rows_in_tile = min(128, p.cout_valid - tbase);
        float* xf = reinterpret_cast<float*>(xpose) + (warp - 2) * (32 * 33);
        float* pbase = reinterpret_cast<float*>(p.out);
#pragma unroll 1
        for (int ci = 0; ci < 8; ++ci) {
          float v[32];
          tmem_ld_32x32(taddr + ci * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) xf[lane * 33 + j] = v[j];
          __syncwarp();
          const long long pos0 = (long long)(2 * pm + (ci >> 2)) * 128 + (ci & 3) * 32 + lane;
          const int nrows = (rows_in_tile - q + 3) >> 2;  // this warp owns logical rows q, q+4, q+8, ...
          for (int rr = 0; rr < nrows; ++rr)
            pbase[(size_t)(tbase + rr * 4 + q) * (size_t)p.sC + pos0] = xf[rr * 33 + lane];
          __syncwarp();
        }
      } else {
#pragma unroll 1
      for (int ci = eg * CHUNKS; ci < (eg + 1) * CHUNKS; ++ci) {
        // this lane's "own" position of the chunk: row r of box (ci >> 2)
        int m = 2 * pm + (ci >> 2);
        const int r = (ci & 3) * 32 + lane;
        const int w = (m % p.tiles_w) * p.bw + r % p.bw;
        m /= p.tiles_w;
        const int h = (m % p.tiles_h) * p.bh + (r / p.bw) % p.bh;
        m /= p.tiles_h;
        const int d = (m % p.tiles_d) * p.bd + r / (p.bw * p.bh);
        const int nb = m / p.tiles_d;
        const bool valid = (r < p.rows_valid) && (w < p.W) && (h < p.H) && (d < p.D);
        const long long roff = p.cls_off[cls] + nb * p.sN + d * p.sD + h * p.sH + w * p.sW;
        const uint32_t vmask = __ballot_sync(0xffffffffu, valid);
        if (p.stats && nb != cur_nb) {  // nb is uniform over the tile (pairs never straddle samples)
          if (cur_nb >= 0) flush();
          cur_nb = nb;
        }
        float v[32];
        tmem_ld_32x32(taddr + ci * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += bias_c;
        if (p.stats) {
          float s = 0.f, ss = 0.f;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float x = ((vmask >> j) & 1u) ? v[j] : 0.f;
            s += x;
            ss += x * x;
          }
          for (int o = 1; o < p.cpg; o <<= 1) {  // a group is cpg adjacent lanes (cpg <= 32)
            s += __shfl_xor_sync(0xffffffffu, s, o);
            ss += __shfl_xor_sync(0xffffffffu, ss, o);
          }
          // a group is owned by ONE lane of ONE warp (cpg <= 32 adjacent channels): plain read-modify-write in
          // program order, so the CTA's partial sums are built in the same order in every run
          if ((lane & (p.cpg - 1)) == 0) {
            gstat[(ch / p.cpg) * 2] += s;
            gstat[(ch / p.cpg) * 2 + 1] += ss;
          }
        }
        // 32 channels x 32 positions -> [position][channel] through shared memory, then 16-byte row stores
#pragma unroll
        for (int j = 0; j < 32; ++j) xp[j * 32 + lane] = __float2half_rn(operand_round(v[j]));
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int pl = k * 8 + (lane >> 2);
          const int cs = (lane & 3) * 8;
          const uint4 val = *reinterpret_cast<const uint4*>(xp + pl * 32 + cs);
          const long long off = __shfl_sync(0xffffffffu, roff, pl);
          if ((vmask >> pl) & 1u) *reinterpret_cast<uint4*>(outp + off + q * 32 + cs) = val;
        }
        __syncwarp();
      }
      }  // MODE
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
    }
    if (p.stats && cur_nb >= 0) flush();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::TM_COLS);
}

}  // namespace b2v
