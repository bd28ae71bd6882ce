// ppm_kernels.cuh — progressive photon mapping:
//   eye pass        Scene::eye_trace            PPM/src/Scene.cpp:286-362
//   hash grid       Scene::build_hash_grid      PPM/src/Scene.cpp:53-93, hash PPM/include/Scene.h:65-68
//   photon pass     Point_light::generate_photon PPM/src/Point_light.cpp:7-30,
//                   Scene::photon_trace          PPM/src/Scene.cpp:106-249, sample_hemisphere :15-44
//   density         Scene::density_estimation    PPM/src/Scene.cpp:363-372
// Both passes are wavefronts over the same vertex pool and the same traversal
// kernel (k_trace_fast<false,7>) as the image renderers.  See include/b200rt.h
// for the lock-free restatement of the per-photon radius / flux update.
#pragma once
#include "rt_shade7.cuh"

#define PPM_ALPHA 0.7f
#define PPM_RTAB (1 << 20)
#define PPM_QCAP 12  // deferred deposit candidates per lane
#ifndef PPM_MIN_BLOCKS
#define PPM_MIN_BLOCKS 5  // 102 registers: room for two rounds of scan loads per lane (variants measured in profiles/r01j_tuning.txt)
#endif
#ifndef PPM_SORT_BITS
#define PPM_SORT_BITS 4  // direction bins (2^bits) the photons of a batch are grouped into before tracing
#endif
#ifndef PPM_SCAN_LOADS
#define PPM_SCAN_LOADS 4  // grid entries each lane fetches per scan step (32 x this many per warp)
#endif

struct HitPoints {
  uint32_t capacity;
  unsigned int* count;
  float4* pos;   // position xyz
  float4* nrm;   // normal xyz, as_float(material)
  float4* wo;    // w_o xyz, as_float(pixel)
  float4* att;   // attenuation rgb, pixel_weight
  unsigned int* n; // accepted-photon count: a dense 4-byte array (8.7 MB at 1080p) that stays L2-resident, so the
                   // radius test r2(n) of a scanned entry does not cost a DRAM round trip
  float4* state;   // S = sum f*phi/R(k) rgb
  unsigned int* n_base;  // multi-GPU: n as of the last synchronisation of the counts (multi_gpu.cuh)
  unsigned int* delta;   // multi-GPU: photons this GPU accepted since then, summed over the GPUs in place, then applied
  // multi-GPU: this GPU is share `part` of `parts`.  Between two synchronisations it claims the acceptance indices
  // n_base + part, n_base + part + parts, ... — every parts-th one, because the other GPUs accept photons of the same
  // stream at the same rate — so the radius a photon is tested against follows the GLOBAL count, not this GPU's
  // fraction of it.  parts = 1 (one GPU): indices 0, 1, 2, ... as the reference's single array.
  uint32_t part, parts;
  float4* dep;   // 4 per hit point, 64 bytes: what a deposit reads — {normal, phong exponent}, {w_o, att.r},
                 // {kd, att.g}, {ks, att.b} (built once per eye pass by k_ppm_grid)
};
__device__ __forceinline__ unsigned int* hp_n(const HitPoints& HP, unsigned int i) { return HP.n + i; }
__device__ __forceinline__ unsigned int hp_stride(const HitPoints& HP) { return HP.parts > 1u ? HP.parts : 1u; }
// accepted photons of hit point i as far as this GPU knows: all GPUs' up to the last sync + its own since
__device__ __forceinline__ unsigned int hp_count(const HitPoints& HP, unsigned int i) {
  if (HP.parts <= 1u) return HP.n[i];
  const unsigned int b = HP.n_base[i];
  return b + (HP.n[i] - b - HP.part) / HP.parts;
}
struct PpmGrid {
  uint32_t num_hash;
  unsigned int* cell_count;  // num_hash + 1 (becomes the fill cursor)
  unsigned int* cell_start;  // num_hash + 1
  unsigned int* items;       // unused by the photon pass (kept for tests / debugging)
  float4* cell_pos;          // per grid entry, in cell order: hit-point position xyz, as_float(hit-point index) —
                             // a photon streams its cell's entries as contiguous 16-byte records
  v3 bbox_min;
  float hash_scale, r0, r0sq;
  const float* rtab;         // R(n) for n < PPM_RTAB
  // The SCAN grid of the photon pass (b200rt_api.cu: ppm_rebuild_scan_grid).  The reference keeps the grid it built
  // from the initial radius for the whole run (cells 2 r0 wide, PPM/src/Scene.cpp:53-93), so a photon goes on testing
  // the ~144 hit points of its cell when the radii have shrunk to a third and two or three of them can still accept
  // it.  Which pairs ACCEPT depends only on |x - q|^2 <= r2(n) and the normals; the reference's cell structure merely
  // finds a superset of them (a photon within r(n) <= r0 of a hit point always lies in a cell of the hit point's
  // r0-box).  Any other superset finds the same pairs, so the device rebuilds the grid as the radii shrink:
  // adaptive != 0: cells are 1 / hash_scale wide (about twice the typical current radius) and a hit point is entered
  // into the cells its CURRENT radius reaches.  Radii only shrink, so a grid stays valid until the next rebuild.
  int adaptive;
};
struct PpmCounters {
  unsigned long long photon_rays, cell_scans, deposits;
  unsigned long long entries_scanned;   // grid entries the photons actually tested (== cell_scans on the reference's grid)
  unsigned int bbox[6];      // ordered-int min xyz / max xyz of the hit-point positions
  unsigned int grid_overflow;
};

__device__ __forceinline__ int ppm_material_type(const b200rt_material& m) {  // PPM/src/Material.cpp:55-61
  if (!is_zero(ld3(m.mirror))) return 1;
  if (!is_zero(ld3(m.transparency))) return 2;
  return 0;
}
__device__ __forceinline__ unsigned int f2ord(float f) {
  const unsigned int u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float ord2f_bits(unsigned int o) {
  const unsigned int u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  float f;
  memcpy(&f, &u, 4);
  return f;
#endif
}
// R(n) = prod_{k<n} (k*a + a)/(k*a + 1): table (the reference's float product) or, beyond it, the closed form
__device__ __forceinline__ float ppm_R(const PpmGrid& G, unsigned int n) {
  if (n < PPM_RTAB) return __ldg(G.rtab + n);
  const double ia = 1.0 / (double)PPM_ALPHA;
  return (float)exp(lgamma((double)n + 1.0) + lgamma(ia) - lgamma((double)n + ia));
}
__device__ __forceinline__ uint32_t ppm_hash(uint32_t num_hash, int ix, int iy, int iz) {
  const uint32_t h = ((uint32_t)ix * 73856093u) ^ ((uint32_t)iy * 19349663u) ^ ((uint32_t)iz * 83492791u);
  return h % num_hash;
}

// ---- eye pass ------------------------------------------------------------------
__global__ void __launch_bounds__(RT_BLOCK) k_ppm_eye_shade(DScene S, Pool P, WaveInfo* waves, int d, HitPoints HP,
                                                            Counters* C, PpmCounters* PC) {
  const uint32_t begin = waves[d].begin, count = waves[d].count;
  const uint32_t next_begin = waves[d + 1].begin;
  const uint32_t stride = gridDim.x * blockDim.x;
  const uint32_t rounds = (count + stride - 1) / stride;
  for (uint32_t r = 0; r < rounds; r++) {
    const uint32_t k = r * stride + blockIdx.x * blockDim.x + threadIdx.x;
    const bool act = k < count;
    const uint32_t slot = begin + (act ? k : 0);
    const float4 ha = P.hit_a[slot];
    const uint32_t top_ref = __float_as_uint(ha.w);
    const bool hit = act && B200RT_REF_KIND(top_ref) != B200RT_KIND_NONE;
    bool want_hp = false, want_r = false, want_t = false;
    v3 p = mk3(0, 0, 0), n = mk3(0, 0, 0), w_o = mk3(0, 0, 0), att = mk3(1, 1, 1);
    v3 dir_r = mk3(0, 0, 0), dir_t = mk3(0, 0, 0), att_r = mk3(0, 0, 0), att_t = mk3(0, 0, 0);
    int mat = 0;
    uint32_t depth = 0;
    if (hit) {
      const float4 ro = P.ray_o[slot], rd = P.ray_d[slot];
      const v3 o = mk3(ro.x, ro.y, ro.z), dir = mk3(rd.x, rd.y, rd.z);
      depth = RT_META_DEPTH(__float_as_uint(rd.w));
      if (depth > 0) { const float4 a = P.rec0[slot]; att = mk3(a.x, a.y, a.z); }
      const Surf7 sf = surface7(S, top_ref, P.hit_b[slot].y, ha.y, ha.z, o, dir, 0.0f, ha.x);
      n = sf.n; mat = sf.material;
      p = o + (dir * ha.x);
      const b200rt_material& m = S.mats[mat];
      const int type = ppm_material_type(m);
      w_o = normalize(o - p);
      if (type == 0) {
        want_hp = true;
      } else if ((int)depth >= S.max_depth) {
        // dropped — PPM/src/Scene.cpp:305-306
      } else {
        const v3 w_r = normalize((n * (2.0f * dot(n, w_o))) - w_o);
        dir_r = w_r;
        if (type == 1) {
          want_r = true; att_r = ld3(m.mirror) * att;
        } else {
          const v3 nl = dot(n, dir) < 0.0f ? n : n * -1.0f;
          const bool into = dot(n, nl) > 0.0f;
          const float air = 1.0f;
          const float nnt = into ? air / m.refraction_index : m.refraction_index / air;
          const float ddn = dot(dir, nl);
          const float cos2t = 1 - nnt * nnt * (1 - ddn * ddn);
          if (cos2t < 0.0f) {
            want_r = true; att_r = ld3(m.transparency) * att;
          } else {
            dir_t = normalize((dir * nnt) - (n * ((into ? 1 : -1) * (ddn * nnt + sqrtf(cos2t)))));
            const float a = m.refraction_index - air, b = m.refraction_index + air;
            const float R0 = a * a / (b * b);
            const float ca = into ? -ddn : dot(dir_t, n);
            const float c = 1 - ca;
            const float fres = R0 + (1 - R0) * c * c * c * c * c;
            const v3 ac = ld3(m.transparency) * att;
            want_t = true;
            if (into) {
              want_r = true; att_r = att * fres;
              att_t = ac * (1.0f - fres);
            } else {
              att_t = ac;
            }
          }
        }
      }
    }
    const uint32_t hi = warp_append(HP.count, want_hp);
    if (want_hp) {
      if (hi < HP.capacity) {
        // pos.w: the vertex's place in its pixel's tree (stratum << 20 | reflect / transmit code) — with the pixel it
        // is the canonical sort key of the hit points (multi_gpu.cuh: k_ppm_order_keys)
        const uint32_t code = depth == 0 ? 0u : (P.key[slot] & 0xfffffu);
        HP.pos[hi] = make_float4(p.x, p.y, p.z, __uint_as_float(((P.sample[slot] & 0xfffu) << 20) | code));
        HP.nrm[hi] = make_float4(n.x, n.y, n.z, __int_as_float(mat));
        HP.wo[hi] = make_float4(w_o.x, w_o.y, w_o.z, __uint_as_float(P.pixel[slot]));
        HP.att[hi] = make_float4(att.x, att.y, att.z, 1.0f);
        atomicMin(&PC->bbox[0], f2ord(p.x)); atomicMin(&PC->bbox[1], f2ord(p.y)); atomicMin(&PC->bbox[2], f2ord(p.z));
        atomicMax(&PC->bbox[3], f2ord(p.x)); atomicMax(&PC->bbox[4], f2ord(p.y)); atomicMax(&PC->bbox[5], f2ord(p.z));
      } else {
        C->overflow = 1;
      }
    }
#pragma unroll
    for (int c = 0; c < 2; c++) {
      const bool want = c == 0 ? want_r : want_t;
      const uint32_t ci = warp_append(&waves[d + 1].count, want);
      if (want) {
        const uint32_t cs = next_begin + ci;
        if (cs < P.capacity) {
          const v3 cd = c == 0 ? dir_r : dir_t;
          const v3 ca = c == 0 ? att_r : att_t;
          const v3 co = p + (cd * S.shadow_eps);
          P.ray_o[cs] = make_float4(co.x, co.y, co.z, 0.0f);
          P.ray_d[cs] = make_float4(cd.x, cd.y, cd.z, __uint_as_float(depth + 1));
          P.rec0[cs] = make_float4(ca.x, ca.y, ca.z, 0.f);
          P.pixel[cs] = P.pixel[slot];
          P.sample[cs] = P.sample[slot];
          // tree code of the child: parent * 3 + (1 reflect | 2 transmit); unique down to depth 12 (3^12 < 2^20)
          P.key[cs] = ((depth == 0 ? 0u : P.key[slot]) * 3u + (uint32_t)(c + 1)) & 0xfffffu;
        } else {
          C->overflow = 1;
        }
      }
    }
  }
}

// ---- hash grid -------------------------------------------------------------------
// mode 0: count entries per cell; mode 1: fill (cell_count is the running cursor, starting at cell_start)
// mode 2: count for a rebuild of the scan grid (the hit points' state is NOT reset)
__global__ void k_ppm_grid(HitPoints HP, PpmGrid G, uint32_t n_hp, int mode, uint32_t items_capacity, PpmCounters* PC,
                           const b200rt_material* mats) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_hp; i += gridDim.x * blockDim.x) {
    const float4 q = HP.pos[i];
    if (mode == 0) {
      HP.state[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      HP.n[i] = HP.parts > 1u ? HP.part : 0u; HP.n_base[i] = 0u;
      const float4 n4 = HP.nrm[i], w4 = HP.wo[i], a4 = HP.att[i];
      const b200rt_material& m = mats[__float_as_int(n4.w)];
      HP.dep[4 * (size_t)i + 0] = make_float4(n4.x, n4.y, n4.z, m.phong_exponent);
      HP.dep[4 * (size_t)i + 1] = make_float4(w4.x, w4.y, w4.z, a4.x);
      HP.dep[4 * (size_t)i + 2] = make_float4(m.diffuse[0], m.diffuse[1], m.diffuse[2], a4.y);
      HP.dep[4 * (size_t)i + 3] = make_float4(m.specular[0], m.specular[1], m.specular[2], a4.z);
    }
    // reach of the hit point: the initial radius (the reference's grid), or its current radius with a margin that
    // covers the rounding of the square root and of the cell arithmetic
    const float rb = G.adaptive ? sqrtf(G.r0sq * ppm_R(G, HP.n[i])) * 1.0001f : G.r0;
    const v3 lo = (mk3(q.x - rb, q.y - rb, q.z - rb) - G.bbox_min) * G.hash_scale;
    const v3 hi = (mk3(q.x + rb, q.y + rb, q.z + rb) - G.bbox_min) * G.hash_scale;
    for (int iz = abs((int)lo.z); iz <= abs((int)hi.z); iz++)
      for (int iy = abs((int)lo.y); iy <= abs((int)hi.y); iy++)
        for (int ix = abs((int)lo.x); ix <= abs((int)hi.x); ix++) {
          const uint32_t hv = ppm_hash(G.num_hash, ix, iy, iz);
          const unsigned int at = atomicAdd(&G.cell_count[hv], 1u);
          if (mode == 1) {
            if (at < items_capacity) {
              if (G.items) G.items[at] = i;
              G.cell_pos[at] = make_float4(q.x, q.y, q.z, __uint_as_float(i));
            } else {
              PC->grid_overflow = 1;
            }
          }
        }
  }
}

// sum over the hit points of R(n) = r2 / r0^2 (double): how far the radii have shrunk since the grid was built
__global__ void k_ppm_sum_R(HitPoints HP, PpmGrid G, uint32_t n_hp, double* out) {
  double acc = 0.0;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_hp; i += gridDim.x * blockDim.x) acc += (double)ppm_R(G, HP.n[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane_id() == 0) atomicAdd(out, acc);
}

// Home cell of every hit point (the cell its position hashes to): second, stable sort key of the canonical numbering,
// so that the hit points a photon meets while scanning one cell sit in a few runs of consecutive indices.
__global__ void k_ppm_home_cell_keys(const float4* pos, PpmGrid G, uint32_t n_hp, uint32_t* keys, uint32_t* idx) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_hp; i += gridDim.x * blockDim.x) {
    const float4 q = pos[i];
    const v3 hh = (mk3(q.x, q.y, q.z) - G.bbox_min) * G.hash_scale;
    keys[i] = ppm_hash(G.num_hash, abs((int)hh.x), abs((int)hh.y), abs((int)hh.z));
    idx[i] = i;
  }
}

// ---- photon pass -------------------------------------------------------------------
// RNG: key of photon i = mix(mix(s0 ^ 'PHOT', lo32(i)), hi32(i)); dims: emission 0,1; bounce b: 8b+2,3 hemisphere,
// 8b+4 roulette, 8b+5 glass roulette (shared with oracle.c).
// Direction of photon i — Point_light::generate_photon, PPM/src/Point_light.cpp:7-30 (theta = 2 pi xi_2: sic,
// SURVEY Appendix B.6).  RNG: key = mix(mix(s0 ^ 'PHOT', lo), hi), dims 0, 1.
__device__ __forceinline__ v3 ppm_photon_dir(uint32_t s0, unsigned long long i, uint32_t* key_out) {
  const uint32_t key = rng_mix(rng_mix(s0 ^ 0x50484F54u, (uint32_t)i), (uint32_t)(i >> 32));
  const float phi = (float)(2 * RT_PI * (double)rng_u01(key, 0));
  const float theta = (float)(2 * RT_PI * (double)rng_u01(key, 1));
  *key_out = key;
  return onb_dir7(mk3(0.0f, 1.0f, 0.0f), theta, phi);
}
// The photons of a batch are a SET (first .. first+n-1): the order they are traced in is ours.  Tracing them sorted
// by direction (16-bit Morton code of the octahedral map) makes consecutive photons hit neighbouring surface points:
// the cell lists a warp streams in the deposit (280 MB of them at 1080p, read at random otherwise: 19 GB of DRAM
// traffic per 10 M photons, profiles/r01j) are then shared through L1/L2, and the first wave's rays are coherent.
__global__ void k_ppm_emit_keys(uint32_t s0, unsigned long long first, uint32_t n, int bits, uint32_t* keys,
                                uint32_t* idx) {
  for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    uint32_t key;
    const v3 d = ppm_photon_dir(s0, first + k, &key);
    const float a = fabsf(d.x) + fabsf(d.y) + fabsf(d.z);
    float u = d.x / a, v = d.z / a;
    if (d.y < 0.0f) {
      const float uu = (1.0f - fabsf(v)) * (u >= 0.0f ? 1.0f : -1.0f), vv = (1.0f - fabsf(u)) * (v >= 0.0f ? 1.0f : -1.0f);
      u = uu; v = vv;
    }
    uint32_t x = (uint32_t)fminf(255.0f, fmaxf(0.0f, (u * 0.5f + 0.5f) * 256.0f));
    uint32_t y = (uint32_t)fminf(255.0f, fmaxf(0.0f, (v * 0.5f + 0.5f) * 256.0f));
    x = (x | (x << 4)) & 0x0F0Fu; x = (x | (x << 2)) & 0x3333u; x = (x | (x << 1)) & 0x5555u;
    y = (y | (y << 4)) & 0x0F0Fu; y = (y | (y << 2)) & 0x3333u; y = (y | (y << 1)) & 0x5555u;
    keys[k] = (x | (y << 1)) >> (16 - bits);
    idx[k] = k;
  }
}
__global__ void k_ppm_emit(DScene S, Pool P, WaveInfo* waves, uint32_t s0, unsigned long long first, uint32_t n,
                           const uint32_t* order) {
  const b200rt_light& L = S.lights[0];
  const v3 flux = ld3(L.intensity) * (float)(RT_PI * 4.0f);
  for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    uint32_t key;
    const v3 d = ppm_photon_dir(s0, first + (order ? order[k] : k), &key);
    P.ray_o[k] = make_float4(L.position[0], L.position[1], L.position[2], 0.0f);
    P.ray_d[k] = make_float4(d.x, d.y, d.z, __uint_as_float(0u));
    P.rec0[k] = make_float4(flux.x, flux.y, flux.z, 0.f);
    P.key[k] = key;
    P.pixel[k] = 0;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) { waves[0].begin = 0; waves[0].count = n; }
}
// ping-pong wave placement: a photon path is a chain, so two half pools suffice
__global__ void k_ppm_next_wave(WaveInfo* waves, int d, uint32_t half, Counters* C) {
  waves[d + 1].begin = ((d + 1) & 1) ? half : 0u;
  waves[d + 1].count = 0;
  C->work_extend = 0;
}
// G: the grid the photons scan (the reference's, or the rebuilt one); ref_start / ref_scale: cell table and cell scale
// of the reference's grid, for the count of hit points ITS loop would have visited (B200PpmStats::cell_scans).
__global__ void __launch_bounds__(RT_BLOCK, PPM_MIN_BLOCKS) k_ppm_photon_shade(DScene S, Pool P, WaveInfo* waves, int d, HitPoints HP,
                                                               PpmGrid G, Counters* C, PpmCounters* PC,
                                                               const unsigned int* ref_start, float ref_scale) {
  const uint32_t begin = waves[d].begin, count = waves[d].count;
  const uint32_t next_begin = waves[d + 1].begin;
  const uint32_t stride = gridDim.x * blockDim.x;
  const uint32_t rounds = (count + stride - 1) / stride;
  unsigned int scans = 0, deposits = 0, scanned = 0;
  for (uint32_t r = 0; r < rounds; r++) {
    const uint32_t k = r * stride + blockIdx.x * blockDim.x + threadIdx.x;
    const bool act = k < count;
    const uint32_t slot = begin + (act ? k : 0);
    const float4 ha = P.hit_a[slot];
    const uint32_t top_ref = __float_as_uint(ha.w);
    const bool hit = act && B200RT_REF_KIND(top_ref) != B200RT_KIND_NONE;
    bool cont = false, dep = false;
    v3 p = mk3(0, 0, 0), ndir = mk3(0, 0, 0), nflux = mk3(0, 0, 0);
    v3 dep_n = mk3(0, 0, 0), dep_flux = mk3(0, 0, 0), dep_wi = mk3(0, 0, 0);
    uint32_t key = 0, b = 0;
    if (hit) {
      const float4 ro = P.ray_o[slot], rd = P.ray_d[slot], fl = P.rec0[slot];
      const v3 o = mk3(ro.x, ro.y, ro.z), dir = mk3(rd.x, rd.y, rd.z);
      const v3 flux = mk3(fl.x, fl.y, fl.z);
      b = RT_META_DEPTH(__float_as_uint(rd.w));
      key = P.key[slot];
      const uint32_t dim = 8u * b;
      const Surf7 sf = surface7(S, top_ref, P.hit_b[slot].y, ha.y, ha.z, o, dir, 0.0f, ha.x);
      const v3 n = sf.n;
      p = o + (dir * ha.x);
      const v3 nl = dot(n, dir) < 0 ? n : n * -1.0f;
      const b200rt_material& m = S.mats[sf.material];
      const int type = ppm_material_type(m);
      if (type == 0) {
        dep = G.num_hash > 0;
        dep_n = n; dep_flux = flux; dep_wi = -normalize(dir);
        // sample_hemisphere(normal, d, probability): cosine — PPM/src/Scene.cpp:15-44
        const float e1 = rng_u01(key, dim + 2), e2 = rng_u01(key, dim + 3);
        const float phi = (float)(2 * RT_PI * (double)e1);
        const float theta = asinf(sqrtf(e2));
        const v3 dd = onb_dir7(n, theta, phi);
        const float prob = (float)((double)fmaxf(0.0f, dot(n, dd)) / RT_PI);
        v3 base = mk3(0, 0, 0);
        const v3 w_i = -normalize(dir);
        const float cti = fmaxf(0.0f, dot(n, w_i));
        if (!(cti > 1.0f || cti <= 0.0f)) {
          const float sct = fmaxf(0.0f, dot(nl, normalize(dd + w_i)));
          base = ld3(m.diffuse) + ((ld3(m.specular) * powf(sct, m.phong_exponent)) / cti);
        }
        const float cto = fmaxf(0.0f, dot(n, dd));
        base = base * cto;
        if (rng_u01(key, dim + 4) < prob) {   // the pdf doubles as survival probability (Appendix B.8)
          cont = true; ndir = dd; nflux = (base * flux) / prob;
        }
      } else {
        const v3 w_o = normalize(o - p);
        const v3 w_r = normalize((n * (2.0f * dot(n, w_o))) - w_o);
        if (type == 1) {
          cont = true; ndir = w_r; nflux = ld3(m.mirror) * flux;
        } else {
          const bool into = dot(n, nl) > 0.0f;
          const float air = 1.0f;
          const float nnt = into ? air / m.refraction_index : m.refraction_index / air;
          const float ddn = dot(dir, nl);
          const float cos2t = 1 - nnt * nnt * (1 - ddn * ddn);
          cont = true; nflux = flux;
          if (cos2t < 0.0f) {
            ndir = w_r;
          } else {
            const v3 rdir = normalize((dir * nnt) - (n * ((into ? 1 : -1) * (ddn * nnt + sqrtf(cos2t)))));
            const float a = m.refraction_index - air, bb = m.refraction_index + air;
            const float R0 = a * a / (bb * bb);
            const float ca = into ? -ddn : dot(rdir, n);
            const float c = 1 - ca;
            const float fres = R0 + (1 - R0) * c * c * c * c * c;
            ndir = (into && rng_u01(key, dim + 5) < fres) ? w_r : rdir;
          }
        }
      }
      // photon_trace: depth++ then `depth >= max_recursion_depth` stops BEFORE the next intersection
      if (cont && (int)(b + 2) >= S.max_depth) cont = false;
    }
    // ---- deposit ------------------------------------------------------------------------------------
    // Scan: the 32 photons of the warp are taken in turn and all lanes scan that photon's cell together
    // (coalesced 16-byte entries, four loads in flight per lane).  An entry that passes the radius test is
    // only QUEUED (hit point, distance, photon lane); the queue is drained afterwards with every lane
    // working on its own candidate, so the dependent DRAM / atomic round trips of an accepted photon
    // (normal -> CAS on n -> material -> vector atomic) overlap across 32 candidates instead of stalling the
    // scan of each photon in turn.
    unsigned int q_hi[PPM_QCAP];
    float q_d2[PPM_QCAP];
    int q_src[PPM_QCAP];
    int qn = 0;
    auto drain = [&]() {
      int qmax = qn;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) qmax = max(qmax, __shfl_xor_sync(0xffffffffu, qmax, o));
      for (int j = 0; j < qmax; j++) {
        const bool has = j < qn;
        const int src = has ? q_src[j] : (int)lane_id();
        const v3 pn = mk3(__shfl_sync(0xffffffffu, dep_n.x, src), __shfl_sync(0xffffffffu, dep_n.y, src),
                          __shfl_sync(0xffffffffu, dep_n.z, src));
        const v3 pf = mk3(__shfl_sync(0xffffffffu, dep_flux.x, src), __shfl_sync(0xffffffffu, dep_flux.y, src),
                          __shfl_sync(0xffffffffu, dep_flux.z, src));
        const v3 w_i = mk3(__shfl_sync(0xffffffffu, dep_wi.x, src), __shfl_sync(0xffffffffu, dep_wi.y, src),
                           __shfl_sync(0xffffffffu, dep_wi.z, src));
        if (!has) continue;
        const unsigned int hi = q_hi[j];
        const float d2 = q_d2[j];
        const float4 hn = __ldg(HP.dep + 4 * (size_t)hi);
        const v3 hnrm = mk3(hn.x, hn.y, hn.z);
        if (!(dot(hnrm, pn) > 1e-3f)) continue;
        // claim index n under the radius the reference would test against: r2(n) = r0^2 * R(n)
        unsigned int n_old = *((volatile unsigned int*)hp_n(HP, hi));
        bool accepted = false;
        for (;;) {
          if (!(d2 <= G.r0sq * ppm_R(G, n_old))) break;
          const unsigned int prev = atomicCAS(hp_n(HP, hi), n_old, n_old + hp_stride(HP));
          if (prev == n_old) { accepted = true; break; }
          n_old = prev;
        }
        if (!accepted) continue;
        deposits++;
        const float cti = dot(hnrm, w_i);
        if (!(cti > 1.0f || cti <= 0.0f)) {
          const float4 hw = __ldg(HP.dep + 4 * (size_t)hi + 1), hd = __ldg(HP.dep + 4 * (size_t)hi + 2),
                       hs = __ldg(HP.dep + 4 * (size_t)hi + 3);
          const float sct = fmaxf(0.0f, dot(hnrm, normalize(mk3(hw.x, hw.y, hw.z) + w_i)));
          const v3 color = (mk3(hd.x, hd.y, hd.z) + ((mk3(hs.x, hs.y, hs.z) * powf(sct, hn.w)) / cti)) *
                           mk3(hw.w, hd.w, hs.w);
          const v3 c = (color * pf) / ppm_R(G, n_old);
          atomicAdd(HP.state + hi, make_float4(c.x, c.y, c.z, 0.0f));   // one 16-byte vector atomic (sm_90+)
        }
      }
      qn = 0;
    };
    // every lane looks its own cell range up first: one memory latency for the 32 photons of the warp
    unsigned int my_c0 = 0, my_c1 = 0;
    if (dep) {
      const v3 hh = (p - G.bbox_min) * G.hash_scale;
      const uint32_t cell = ppm_hash(G.num_hash, abs((int)hh.x), abs((int)hh.y), abs((int)hh.z));
      my_c0 = G.cell_start[cell]; my_c1 = G.cell_start[cell + 1];
      scanned += my_c1 - my_c0;
      if (G.adaptive) {
        const v3 hr = (p - G.bbox_min) * ref_scale;
        const uint32_t rc = ppm_hash(G.num_hash, abs((int)hr.x), abs((int)hr.y), abs((int)hr.z));
        scans += ref_start[rc + 1] - ref_start[rc];
      } else {
        scans += my_c1 - my_c0;
      }
    }
    // The scan is a flat sequence of steps (photon, 32 x PPM_SCAN_LOADS entries of its cell), walked by the whole
    // warp; the entries of step k+1 are requested before step k is consumed, so a lane always has two rounds of
    // 16-byte loads in flight (the kernel waits on memory: long-scoreboard 58 % of its stall samples, profiles/r01j).
    unsigned dmask = __ballot_sync(0xffffffffu, dep);
    int n_src = -1;                 // photon (lane) the NEXT step belongs to
    unsigned int n_base = 0, n_c1 = 0;
    auto advance = [&]() {          // moves (n_src, n_base) to the following step; n_src = -1 when there is none
      if (n_src >= 0) {
        n_base += 32u * PPM_SCAN_LOADS;
        if (n_base < n_c1) return;
      }
      for (;;) {
        if (!dmask) { n_src = -1; return; }
        n_src = __ffs(dmask) - 1;
        dmask &= dmask - 1;
        n_base = __shfl_sync(0xffffffffu, my_c0, n_src);
        n_c1 = __shfl_sync(0xffffffffu, my_c1, n_src);
        if (n_base < n_c1) return;
      }
    };
    float4 nxt[PPM_SCAN_LOADS];
    auto fetch = [&]() {            // streaming loads (evict-first): 280 MB of entries, each read once per photon
#pragma unroll
      for (int j = 0; j < PPM_SCAN_LOADS; j++) {
        const unsigned int q = n_base + 32u * j + lane_id();
        nxt[j] = (n_src >= 0 && q < n_c1) ? __ldcs(G.cell_pos + q) : make_float4(RT_INF, RT_INF, RT_INF, 0.f);
      }
    };
    advance();
    fetch();
    while (n_src >= 0) {
      const int src = n_src;
      float4 ent[PPM_SCAN_LOADS];
#pragma unroll
      for (int j = 0; j < PPM_SCAN_LOADS; j++) ent[j] = nxt[j];
      advance();
      fetch();
      const v3 pp = mk3(__shfl_sync(0xffffffffu, p.x, src), __shfl_sync(0xffffffffu, p.y, src), __shfl_sync(0xffffffffu, p.z, src));
      // r2(n) <= r0^2, so an entry farther than r0 cannot be accepted (same decision, fewer loads); for the others
      // the accepted-photon counts are fetched together, THEN tested: one memory round trip per step instead of one
      // per surviving entry
      float d2s[PPM_SCAN_LOADS];
      unsigned int nn[PPM_SCAN_LOADS];
#pragma unroll
      for (int j = 0; j < PPM_SCAN_LOADS; j++) {
        const v3 v = mk3(ent[j].x, ent[j].y, ent[j].z) - pp;
        d2s[j] = dot(v, v);
        nn[j] = 0u;
        if (d2s[j] <= G.r0sq) nn[j] = *((volatile unsigned int*)hp_n(HP, __float_as_uint(ent[j].w)));
      }
      float rr[PPM_SCAN_LOADS];
#pragma unroll
      for (int j = 0; j < PPM_SCAN_LOADS; j++) rr[j] = (d2s[j] <= G.r0sq) ? ppm_R(G, nn[j]) : 0.0f;
#pragma unroll
      for (int j = 0; j < PPM_SCAN_LOADS; j++) {
        // the reference accepts iff (normals agree) && (|v|^2 <= r2); the radius goes first and rejects almost
        // every entry once the radii have shrunk
        if (!(d2s[j] <= G.r0sq) || !(d2s[j] <= G.r0sq * rr[j])) continue;
        q_hi[qn] = __float_as_uint(ent[j].w); q_d2[qn] = d2s[j]; q_src[qn] = src; qn++;
      }
      if (__any_sync(0xffffffffu, qn > PPM_QCAP - PPM_SCAN_LOADS)) drain();
    }
    drain();
    const uint32_t ci = warp_append(&waves[d + 1].count, cont);
    if (cont) {
      const uint32_t cs = next_begin + ci;
      const v3 co = p + (ndir * S.shadow_eps);
      P.ray_o[cs] = make_float4(co.x, co.y, co.z, 0.0f);
      P.ray_d[cs] = make_float4(ndir.x, ndir.y, ndir.z, __uint_as_float(b + 1));
      P.rec0[cs] = make_float4(nflux.x, nflux.y, nflux.z, 0.f);
      P.key[cs] = key;
      P.pixel[cs] = 0;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    scans += __shfl_xor_sync(0xffffffffu, scans, o);
    deposits += __shfl_xor_sync(0xffffffffu, deposits, o);
    scanned += __shfl_xor_sync(0xffffffffu, scanned, o);
  }
  if (lane_id() == 0) {
    atomicAdd(&PC->cell_scans, (unsigned long long)scans);
    atomicAdd(&PC->deposits, (unsigned long long)deposits);
    atomicAdd(&PC->entries_scanned, (unsigned long long)scanned);
  }
}
__global__ void k_ppm_count_rays(const WaveInfo* waves, int d, PpmCounters* PC) {
  PC->photon_rays += waves[d].count;
}

// ---- density estimation -----------------------------------------------------------------
// packs {S.rgb, n} per hit point (the additive state; all-reduced across GPUs)
__global__ void k_ppm_export(HitPoints HP, uint32_t n_hp, float4* out) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_hp; i += gridDim.x * blockDim.x) {
    const float4 s = HP.state[i];
    out[i] = make_float4(s.x, s.y, s.z, (float)hp_count(HP, i));
  }
}
// tau/(pi r2 N) = S/(pi r0^2 N) since tau = R(n) S and r2 = r0^2 R(n)
__global__ void k_ppm_density(HitPoints HP, uint32_t n_hp, const float4* S_n, float r0sq, double total, float4* accum) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_hp; i += gridDim.x * blockDim.x) {
    const float4 s = S_n[i];
    const float4 w = HP.wo[i], a = HP.att[i];
    const uint32_t pix = __float_as_uint(w.w);
    const float kk = (float)(1.0f / (RT_PI * (double)r0sq * total));
    const v3 col = mk3(s.x, s.y, s.z) * kk;
    float* px = reinterpret_cast<float*>(accum + pix);
    atomicAdd(px + 0, col.x * a.w); atomicAdd(px + 1, col.y * a.w); atomicAdd(px + 2, col.z * a.w);
    atomicAdd(px + 3, a.w);
  }
}
__global__ void k_ppm_pack_hit_points(HitPoints HP, uint32_t n_hp, float* out16) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_hp; i += gridDim.x * blockDim.x) {
    const float4 p = HP.pos[i], n = HP.nrm[i], w = HP.wo[i], a = HP.att[i];
    float* o = out16 + 16 * (size_t)i;
    o[0] = p.x; o[1] = p.y; o[2] = p.z; o[3] = n.x; o[4] = n.y; o[5] = n.z; o[6] = w.x; o[7] = w.y; o[8] = w.z;
    o[9] = a.x; o[10] = a.y; o[11] = a.z; o[12] = (float)__float_as_int(n.w); o[13] = (float)__float_as_uint(w.w);
    o[14] = a.w; o[15] = (float)hp_count(HP, i);
  }
}
