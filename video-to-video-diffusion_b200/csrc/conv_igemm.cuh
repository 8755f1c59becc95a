// Implicit-GEMM Conv3d / ConvTranspose3d for sm_100a.
//
// Replaces the cuDNN/oneDNN convolutions the reference reaches through nn.Conv3d / nn.ConvTranspose3d
// (reference models/unet3d.py:56,96,102,204,218,257,331 and models/vae.py:27,45,65,86,134,137,161,188).
//
//   activations : NDHWC fp16 ("cl16"), channels padded to a multiple of 64
//   weights     : [tap][Cout_pad][Cin_total] fp16 (K-major rows for the B operand)
//   M tile      : a (bd x bh x bw) box of <=128 output positions of one sample, fetched per filter tap with
//                 one 5-D TMA box load whose start coordinate is shifted by the tap offset; out-of-bounds
//                 elements are zero-filled by TMA, which IS the conv zero padding
//   K loop      : taps x 64-channel chunks (x up to two sources: the U-Net skip concat is never materialised)
//   MMA         : tcgen05.mma kind::f16, M=128 x N=BN x K=16, fp32 accumulators double-buffered in TMEM
//   epilogue    : +bias, optional per-(sample,group) sum / sum-of-squares for the following GroupNorm,
//                 optional tanh, store fp16 NDHWC (vectorised) or fp32 with arbitrary strides (NCDHW heads,
//                 sub-pixel interleave of the transposed conv)
//
// Strided (1,2,2) convs read four parity views of the input (four tensor maps with doubled strides);
// transposed (1,2,2) convs run as four output-parity classes of 3x2x2 taps each.
#pragma once
#include "conv_params.h"
#include "ptx.cuh"

namespace b2v {

// PAIR: two CTAs of a (2,1,1) cluster run one M = 256 x N = BN MMA (tcgen05 cta_group::2): each CTA stages its own
// 128-position A box and HALF of the weight tile, so the weight bytes moved L2 -> shared memory and read from shared
// memory per FLOP are halved (the step is power-limited: operand movement is what is left to save).
template <int BN, bool PAIR = false>
struct ConvCfg {
  static constexpr int A_BYTES = 128 * 128;
  static constexpr int B_BYTES = PAIR ? BN * 64 : BN * 128;
  static constexpr int STAGE = A_BYTES + B_BYTES;
  static constexpr int NSTAGE = PAIR ? 6 : ((BN == 256) ? 4 : (BN == 128 ? 6 : 8));
  static constexpr int TM_COLS = (2 * BN < 32) ? 32 : 2 * BN;
  // epilogue warps: two per TMEM lane quarter for the wide tiles (each takes half of the tile's columns), so the one
  // epilogue nothing overlaps -- the last tile's -- and the short-K layers' epilogues take half as long
  static constexpr int EW = (BN >= 64) ? 8 : 4;
  static constexpr int THREADS = 64 + 32 * EW;
  static constexpr int SMEM = NSTAGE * STAGE + 1024 /*align slack*/ + 256 /*barriers*/ + 1024 * EW /*group stats, per warp*/;
};

// Sum N per-lane values across the 32 lanes of a warp with N-1 + log2(32/N) shuffles (instead of 5 per value):
// at every stage half of the values travel to the partner lane.  Returns the total of value `idx`
// = top log2(N) bits of the lane id; all lanes of one idx class hold the same result.
template <int N>
__device__ __forceinline__ float multi_reduce(float (&v)[N], int lane) {
  constexpr int LOG = (N == 32) ? 5 : (N == 16) ? 4 : (N == 8) ? 3 : (N == 4) ? 2 : 1;
#pragma unroll
  for (int s = 0; s < LOG; ++s) {
    const int off = 16 >> s;
    const int n = N >> s;
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int j = 0; j < n / 2; ++j) {
      const float send = up ? v[j] : v[j + n / 2];
      const float keep = up ? v[j + n / 2] : v[j];
      v[j] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
#pragma unroll
  for (int off = (16 >> LOG); off > 0; off >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], off);
  return v[0];
}

// per-(sample, group) sum / sum-of-squares of one 32-column chunk, accumulated into THIS WARP's shared-memory
// table (the four tables are folded in a fixed order and flushed to global memory with one fixed-point atomic per
// group when the CTA moves to another sample / n-tile).  Every table entry is only ever updated by one lane of its
// warp, in program order, so the accumulation order -- and with it the fp32 result -- is the same in every run.
template <int SUB>  // columns per group inside the chunk: min(channels-per-group, 32)
__device__ __forceinline__ void chunk_stats(const float* v, bool valid, float* sstat, int glocal, int lane) {
  constexpr int SEGS = 32 / SUB, N = 2 * SEGS, LOG = (N == 32) ? 5 : (N == 16) ? 4 : (N == 8) ? 3 : (N == 4) ? 2 : 1;
  float acc[N];
#pragma unroll
  for (int seg = 0; seg < SEGS; ++seg) {
    float s = 0.f, ss = 0.f;
#pragma unroll
    for (int j = 0; j < SUB; ++j) {
      const float x = valid ? v[seg * SUB + j] : 0.f;
      s += x;
      ss += x * x;
    }
    acc[2 * seg] = s;
    acc[2 * seg + 1] = ss;
  }
  const float tot = multi_reduce<N>(acc, lane);
  if ((lane & ((32 >> LOG) - 1)) == 0) sstat[glocal * 2 + (lane >> (5 - LOG))] += tot;
}

// Split-K for small-M layers (few output tiles, e.g. the 6x6 level at batch 1: 28 tiles on 148 SMs): work unit =
// (tile, k-split); every unit accumulates its slice of the (tap, chunk) loop and stores its fp32 partial tile into
// ITS OWN slab of the workspace (plain vector stores); splitk_finalize_kernel sums the slabs in a fixed order and
// applies bias, statistics and the fp16 conversion -- deterministic, unlike atomics.  splitk == 1 is the ordinary path.
__device__ __forceinline__ int unit_k0(int unit, int ksteps, int S) { return (int)((long long)(unit % S) * ksteps / S); }
__device__ __forceinline__ int unit_k1(int unit, int ksteps, int S) {
  return (int)((long long)(unit % S + 1) * ksteps / S);
}

template <int BN, bool PAIR = false>
__global__ void __launch_bounds__(ConvCfg<BN, PAIR>::THREADS, 1) conv_igemm_kernel(const __grid_constant__ ConvParams p) {
  using Cfg = ConvCfg<BN, PAIR>;
  constexpr int NSTAGE = Cfg::NSTAGE;
  constexpr int EW = Cfg::EW;
  constexpr int CH = (BN >= 32) ? 32 : 16;  // epilogue column chunk

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + NSTAGE * Cfg::STAGE);
  uint64_t* empty = full + NSTAGE;
  uint64_t* tfull = empty + NSTAGE;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  float* sstat = reinterpret_cast<float*>(smem + NSTAGE * Cfg::STAGE + 256);  // [EW epilogue warps][<=128 groups][2]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // PAIR: a work unit is two consecutive m-tiles (one per CTA) of one (class, n-tile); both CTAs of a cluster walk
  // the same unit sequence.  m_units = m-tiles (or m-tile pairs) per (class, n-tile).
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const int m_tiles = p.tiles_w * p.tiles_h * p.tiles_d * p.batch;
  const int m_units = PAIR ? (m_tiles + 1) / 2 : m_tiles;
  const int total_tiles = m_units * p.nclass * p.n_tiles;
  const int total_units = total_tiles * p.splitk;
  const int chunks = p.src_chunks0 + p.src_chunks1;
  const int unit0 = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int unit_step = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&p.tmA[0]);
    tma_prefetch_desc(&p.tmB);
    for (int i = 0; i < NSTAGE; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], PAIR ? 2 * EW : EW);  // PAIR: the epilogue warps of both CTAs release the leader's accumulator
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    if constexpr (PAIR) tmem_alloc_2sm(tmem_slot, Cfg::TM_COLS);
    else tmem_alloc(tmem_slot, Cfg::TM_COLS);
  }
  for (int i = threadIdx.x; i < 256 * EW; i += blockDim.x) sstat[i] = 0.f;
  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all();  // the peer's barriers must be initialised before anything signals them
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // prologue done (barriers, TMEM, descriptors): let the next kernel start its own, then wait for our producer
  pdl_trigger();
  pdl_wait();

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      const uint32_t a_bytes = (uint32_t)p.rows_valid * 128u;
      int stage = 0;
      uint32_t phase = 0;
      const int ksteps = p.ntaps * chunks;
      for (int unit = unit0; unit < total_units; unit += unit_step) {
        const int tile = unit / p.splitk;
        int m = PAIR ? 2 * (tile % m_units) + (int)rank : tile % m_units;
        int rest = tile / m_units;
        const int cls = rest % p.nclass;
        const int n0 = (rest / p.nclass) * BN;
        const int w0 = (m % p.tiles_w) * p.bw;
        m /= p.tiles_w;
        const int h0 = (m % p.tiles_h) * p.bh;
        m /= p.tiles_h;
        const int d0 = (m % p.tiles_d) * p.bd;
        const int nb = m / p.tiles_d;  // == batch for the missing second tile of an odd count: TMA zero-fills it
        const int k0 = unit_k0(unit, ksteps, p.splitk), k1 = unit_k1(unit, ksteps, p.splitk);
        int t = k0 / chunks, c = k0 % chunks;
        for (int k = k0; k < k1; ++k) {
          const int tg = cls * p.ntaps + t;
          const int32_t tp = p.taps[tg];
          const int map = tp >> 24;
          const int cd = d0 + ((tp >> 16) & 0xff) - 8;
          const int chh = h0 + ((tp >> 8) & 0xff) - 8;
          const int cw = w0 + (tp & 0xff) - 8;
          const int src = (c >= p.src_chunks0) ? 1 : 0;
          const int cc = src ? (c - p.src_chunks0) : c;
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::STAGE;
          if constexpr (PAIR) {
            // the leader's barrier counts the bytes of both CTAs; its own arrive.expect_tx is the one pending arrival
            if (rank == 0) mbar_arrive_expect_tx(&full[stage], 2u * (a_bytes + (uint32_t)Cfg::B_BYTES));
            const uint32_t lbar = mapa_u32(smem_u32(&full[stage]), 0);
            tma_load_5d_2sm(sa, &p.tmA[map + src], lbar, cc * 64, cw, chh, cd, nb);
            tma_load_3d_2sm(sa + Cfg::A_BYTES, &p.tmB2, lbar, c * 64, n0 + (int)rank * (BN / 2), tg);
          } else {
            mbar_arrive_expect_tx(&full[stage], a_bytes + (uint32_t)Cfg::B_BYTES);
            tma_load_5d(sa, &p.tmA[map + src], &full[stage], cc * 64, cw, chh, cd, nb);
            tma_load_3d(sa + Cfg::A_BYTES, &p.tmB, &full[stage], c * 64, n0, tg);
          }
          if (++stage == NSTAGE) {
            stage = 0;
            phase ^= 1;
          }
          if (++c == chunks) {
            c = 0;
            ++t;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (PAIR: the leader CTA only)
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(PAIR ? 256 : 128, BN, 0);
      const int ksteps = p.ntaps * chunks;
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int unit = unit0; unit < total_units; unit += unit_step, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
        const int nk = unit_k1(unit, ksteps, p.splitk) - unit_k0(unit, ksteps, p.splitk);
        for (int k = 0; k < nk; ++k) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE);
          const uint64_t adesc = umma_desc_sw128(sa);
          const uint64_t bdesc = umma_desc_sw128(sa + Cfg::A_BYTES);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            if constexpr (PAIR)
              umma_f16_2sm(tmem_d, adesc + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), idesc, (k | kk) ? 1u : 0u);
            else
              umma_f16(tmem_d, adesc + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), idesc, (k | kk) ? 1u : 0u);
          }
          if constexpr (PAIR) umma_commit_2sm(&empty[stage], 3);  // frees the stage in both CTAs
          else umma_commit(&empty[stage]);
          if (++stage == NSTAGE) {
            stage = 0;
            phase ^= 1;
          }
        }
        if constexpr (PAIR) umma_commit_2sm(&tfull[acc], 3);
        else umma_commit(&tfull[acc]);
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..2+EW-1)
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    constexpr int EG = EW / 4;                    // warps per lane quarter
    const int eg = (warp - 2) >> 2;               // which share of the tile's columns this warp handles
    constexpr int COLS = (BN / EG < CH) ? CH : BN / EG;
    const int r = q * 32 + lane;
    const int rw = r % p.bw;
    const int rh = (r / p.bw) % p.bh;
    const int rd = r / (p.bw * p.bh);
    const int et = threadIdx.x - 64;           // 0..32*EW-1 within the epilogue warps
    const int ng = (BN + p.cpg - 1) / p.cpg;   // groups touched by one n-tile
    int cur_key = -1, cur_nb = 0, cur_g0 = 0;
    float* wstat = sstat + (warp - 2) * 256;  // this warp's table
    // flush the CTA-local group statistics: one global fixed-point atomic per (group, moment)
    auto flush = [&]() {
      asm volatile("bar.sync 1, %0;" ::"n"(32 * EW) : "memory");
      if (et < 2 * ng) {
        float val = 0.f;
#pragma unroll
        for (int wi = 0; wi < EW; ++wi) {  // the warps' tables in index order: fixed summation order
          val += sstat[wi * 256 + et];
          sstat[wi * 256 + et] = 0.f;
        }
        const int g = cur_g0 + (et >> 1);
        if (g < p.groups) stat_add(p.stats + ((size_t)cur_nb * p.groups + g) * 2 + (et & 1), val);
      }
      asm volatile("bar.sync 1, %0;" ::"n"(32 * EW) : "memory");
    };
    int it = 0;
    const uint32_t tempty_leader = PAIR ? mapa_u32(smem_u32(&tempty[0]), 0) : 0u;
    for (int unit = unit0; unit < total_units; unit += unit_step, ++it) {
      const int tile = unit / p.splitk;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      int m = PAIR ? 2 * (tile % m_units) + (int)rank : tile % m_units;
      if (PAIR && m >= m_tiles) {  // odd tile count: this CTA's half of the last pair is padding
        mbar_wait(&tfull[acc], acc_phase);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(tempty_leader + (uint32_t)acc * 8u);
        continue;
      }
      int rest = tile / m_units;
      const int cls = rest % p.nclass;
      const int n0 = (rest / p.nclass) * BN;
      const int w = (m % p.tiles_w) * p.bw + rw;
      m /= p.tiles_w;
      const int h = (m % p.tiles_h) * p.bh + rh;
      m /= p.tiles_h;
      const int d = (m % p.tiles_d) * p.bd + rd;
      const int nb = m / p.tiles_d;
      const bool valid = (r < p.rows_valid) && (w < p.W) && (h < p.H) && (d < p.D);
      const long long roff = p.cls_off[cls] + nb * p.sN + d * p.sD + h * p.sH + w * p.sW;
      if (p.stats && p.splitk == 1) {
        const int key = nb * p.n_tiles + n0 / BN;
        if (key != cur_key) {
          if (cur_key >= 0) flush();
          cur_key = key;
          cur_nb = nb;
          cur_g0 = n0 / p.cpg;
        }
      }

      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
#pragma unroll 1
      for (int c0 = eg * COLS; c0 < (eg + 1) * COLS && c0 < BN; c0 += CH) {
        float v[CH];
        if constexpr (CH == 32)
          tmem_ld_32x32(taddr + c0, v);
        else
          tmem_ld_32x16(taddr + c0, v);
        tmem_ld_wait();
        const int cg = n0 + c0;
        if (cg >= p.cout_valid) break;
        if (p.splitk > 1) {  // partial tile -> this k-split's slab of the workspace (same NDHWC indexing as the output)
          if (valid) {
            float4* wsp = reinterpret_cast<float4*>(p.ws + (long long)(unit % p.splitk) * p.ws_slab + roff + cg);
#pragma unroll
            for (int j = 0; j < CH; j += 4) wsp[j / 4] = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          }
          continue;
        }
        {  // the chunk's biases as 16-byte loads (cg is a multiple of 16: aligned)
          const float4* b4 = reinterpret_cast<const float4*>(p.bias + cg);
#pragma unroll
          for (int j = 0; j < CH; j += 4) {
            const float4 bb = __ldg(b4 + j / 4);
            v[j] += bb.x, v[j + 1] += bb.y, v[j + 2] += bb.z, v[j + 3] += bb.w;
          }
        }
        if constexpr (CH == 32) {
          if (p.stats && p.splitk == 1) {
            const int gl = cg / p.cpg - cur_g0;
            if (p.cpg >= 32) chunk_stats<32>(v, valid, wstat, gl, lane);
            else if (p.cpg == 16) chunk_stats<16>(v, valid, wstat, gl, lane);
            else if (p.cpg == 8) chunk_stats<8>(v, valid, wstat, gl, lane);
            else if (p.cpg == 4) chunk_stats<4>(v, valid, wstat, gl, lane);
            else chunk_stats<2>(v, valid, wstat, gl, lane);
          }
        }
        if (p.act == ACT_TANH) {
#pragma unroll
          for (int j = 0; j < CH; ++j) v[j] = tanhf(v[j]);
        }
        if (valid) {
          if (p.out_mode == OUT_CL16) {
            __half* o = reinterpret_cast<__half*>(p.out) + roff + cg;
#pragma unroll
            for (int j = 0; j < CH; j += 8) {
              __half2 h0 = __floats2half2_rn(operand_round(v[j + 0]), operand_round(v[j + 1]));
              __half2 h1 = __floats2half2_rn(operand_round(v[j + 2]), operand_round(v[j + 3]));
              __half2 h2 = __floats2half2_rn(operand_round(v[j + 4]), operand_round(v[j + 5]));
              __half2 h3 = __floats2half2_rn(operand_round(v[j + 6]), operand_round(v[j + 7]));
              uint4 u;
              u.x = *reinterpret_cast<uint32_t*>(&h0);
              u.y = *reinterpret_cast<uint32_t*>(&h1);
              u.z = *reinterpret_cast<uint32_t*>(&h2);
              u.w = *reinterpret_cast<uint32_t*>(&h3);
              *reinterpret_cast<uint4*>(o + j) = u;
            }
          } else {
            float* o = reinterpret_cast<float*>(p.out) + roff;
#pragma unroll
            for (int j = 0; j < CH; ++j)
              if (cg + j < p.cout_valid) o[(long long)(cg + j) * p.sC] = v[j];
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (PAIR) mbar_arrive_cluster(tempty_leader + (uint32_t)acc * 8u);
        else mbar_arrive(&tempty[acc]);
      }
    }
    if (p.stats && p.splitk == 1 && cur_key >= 0) flush();
  }

  tc_fence_before();
  if constexpr (PAIR) {
    cluster_sync_all();  // both CTAs are done with the shared accumulator and with each other's barriers
    if (warp == 1) tmem_dealloc_2sm(tmem_base, Cfg::TM_COLS);
  } else {
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, Cfg::TM_COLS);
  }
}

}  // namespace b2v
