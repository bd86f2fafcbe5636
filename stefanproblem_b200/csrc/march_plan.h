// Host-side plan of the "march" kernels (ipdg_march.cuh, ldg_march.cuh): which cells the
// marching loop stores, which ones go to the kernels' tail, and how the (strip group,
// column segment) tiles are cut so that every CTA gets the same cost.
//
// Decomposition: a warp owns a strip of kStripCells = 30 cells in y (+1 halo cell each
// side = 32 lanes, lane = j - j0 + 1) and marches along x; a CTA = `group_warps` adjacent
// strips (a strip group) over a column segment.
//
// A cell is "interior-uniform" iff its four face neighbours exist and have its own phase;
// those are the cells the loop stores, with compile-time table offsets.  Every other cell
// (boundary cells, cells next to the other phase) goes to the fix list, which each CTA's
// tail recomputes with the general per-cell row function.
//
// Run list of a strip: {end column, mode, phase mask, store mask}; bit `lane` of the store
// mask = the lane's cell is interior-uniform, of the phase mask = it has phase code 2.
// mode & 3: 0 nothing to store, 1 / 2 every stored cell has phase code 1 / 2, 3 both occur
// (the loop runs both phase bodies); mode & 4: the stored lanes are contiguous.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <climits>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace stefan {

constexpr int kStripCells = 30;

// One entry per CTA of a march kernel: {strip group, first column, one past the last, -}
constexpr int kMaxMarchCtas = 640;
struct SegTab {
  int4 e[kMaxMarchCtas];
};
struct SegCache {
  int cbeg, cend, max_ctas, n;
  SegTab tab;
};

struct MarchPlan {
  int nx = 0, ny = 0, nstrips = 0, ngroups = 0, group_warps = 0;
  std::vector<int> mixed_cum;    // [nstrips][nx+1] running count of phase-mixing columns
  std::vector<int> p2_cum;       // [nstrips][nx+1] running count of columns whose stored cells are (mostly) phase 2
  std::vector<int> fix_col_ofs;  // the fix list is sorted by column: [nx+1] offsets
  std::vector<SegCache> seg_cache;
  int4* runs_dev = nullptr;
  int* run_ofs_dev = nullptr;
  int2* fix_dev = nullptr;  // {cell, code}: code = phase index | tL << 1 | tR << 3 | tD << 5 | tU << 7 (face types)
  int nfix = 0;

  std::vector<int4> runs_host;   // what compute() produced (kept: small, and the tests read it)
  std::vector<int> run_ofs_host;
  std::vector<int32_t> fix_host;
  std::vector<int32_t> fix_code_host;  // per fix cell, see fix_dev

  // ph: padded phase array (nx+2) x (ny+2), 0 = no cell.  Returns a cudaError_t.
  cudaError_t build(const int8_t* ph, int nx_, int ny_, int group_warps_) {
    compute(ph, nx_, ny_, group_warps_);
    const std::vector<int4>& runs = runs_host;
    const std::vector<int>& run_ofs = run_ofs_host;
    const std::vector<int32_t>& fix = fix_host;
    cudaError_t e = cudaMalloc(&runs_dev, runs.size() * sizeof(int4));
    if (e == cudaSuccess) e = cudaMemcpy(runs_dev, runs.data(), runs.size() * sizeof(int4), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMalloc(&run_ofs_dev, run_ofs.size() * sizeof(int));
    if (e == cudaSuccess)
      e = cudaMemcpy(run_ofs_dev, run_ofs.data(), run_ofs.size() * sizeof(int), cudaMemcpyHostToDevice);
    if (e == cudaSuccess && nfix > 0) e = cudaMalloc(&fix_dev, fix.size() * sizeof(int2));
    if (e == cudaSuccess && nfix > 0) {
      std::vector<int2> rec(fix.size());
      for (size_t k = 0; k < fix.size(); ++k) rec[k] = make_int2(fix[k], fix_code_host[k]);
      e = cudaMemcpy(fix_dev, rec.data(), rec.size() * sizeof(int2), cudaMemcpyHostToDevice);
    }
    return e;
  }

  // the host part of build(): masks, run lists, fix list, mixed-column counts (no CUDA calls)
  void compute(const int8_t* ph, int nx_, int ny_, int group_warps_) {
    nx = nx_;
    ny = ny_;
    group_warps = group_warps_;
    nstrips = (ny + kStripCells - 1) / kStripCells;
    ngroups = (nstrips + group_warps - 1) / group_warps;
    const int ld = ny + 2;
    std::vector<int32_t>& fix = fix_host;
    std::vector<int4>& runs = runs_host;
    std::vector<int>& run_ofs = run_ofs_host;
    fix.clear();
    fix_code_host.clear();
    runs.clear();
    run_ofs.assign(nstrips + 1, 0);  // [nstrips] = total: the kernels bisect a strip's list
    seg_cache.clear();
    std::vector<uint32_t> smask((size_t)nstrips * nx), pmask((size_t)nstrips * nx);
    for (int i = 0; i < nx; ++i) {
      const int8_t* col = ph + (size_t)(i + 1) * ld + 1;
      for (int j = 0; j < ny; ++j) {
        const int8_t s = col[j];
        const bool uniform = col[j - 1] == s && col[j + 1] == s && col[j - ld] == s && col[j + ld] == s;
        const int sidx = j / kStripCells, lane = j - sidx * kStripCells + 1;
        if (uniform) {
          smask[(size_t)sidx * nx + i] |= 1u << lane;
          if (s == 2) pmask[(size_t)sidx * nx + i] |= 1u << lane;
        } else {
          fix.push_back((int32_t)((int64_t)i * ny + j));
          // face types as ip_face_type(s, neighbour): 0 same phase, 1 other phase, 2 no cell
          auto ft = [&](int8_t sn) { return sn == 0 ? 2 : (sn == s ? 0 : 1); };
          fix_code_host.push_back((s == 2 ? 1 : 0) | ft(col[j - ld]) << 1 | ft(col[j + ld]) << 3 | ft(col[j - 1]) << 5 |
                                  ft(col[j + 1]) << 7);
        }
      }
    }
    mixed_cum.assign((size_t)nstrips * (nx + 1), 0);
    p2_cum.assign((size_t)nstrips * (nx + 1), 0);
    for (int sidx = 0; sidx < nstrips; ++sidx) {
      run_ofs[sidx] = (int)runs.size();
      const uint32_t* sm = smask.data() + (size_t)sidx * nx;
      const uint32_t* pm = pmask.data() + (size_t)sidx * nx;
      auto mode_of = [&](int i) {
        if (sm[i] == 0) return 0;
        int m = pm[i] == 0 ? 1 : (pm[i] == sm[i] ? 2 : 3);
        const uint32_t t = sm[i] >> __builtin_ctz(sm[i]);
        if ((t & (t + 1)) == 0) m |= 4;
        return m;
      };
      for (int i = 0; i < nx;) {
        int e = i + 1;
        while (e < nx && sm[e] == sm[i] && pm[e] == pm[i]) ++e;
        runs.push_back(make_int4(e, mode_of(i), (int)pm[i], (int)sm[i]));
        i = e;
      }
      for (int k = 0; k < 3; ++k) runs.push_back(make_int4(INT_MAX, 0, 0, 0));  // sentinels
      int* cum = mixed_cum.data() + (size_t)sidx * (nx + 1);
      for (int i = 0; i < nx; ++i) cum[i + 1] = cum[i] + ((sm[i] != 0 && pm[i] != 0 && pm[i] != sm[i]) ? 1 : 0);
      int* c2 = p2_cum.data() + (size_t)sidx * (nx + 1);
      for (int i = 0; i < nx; ++i) c2[i + 1] = c2[i] + (2 * __builtin_popcount(pm[i]) > __builtin_popcount(sm[i]) ? 1 : 0);
    }
    run_ofs[nstrips] = (int)runs.size();
    nfix = (int)fix.size();
    fix_col_ofs.assign(nx + 1, 0);
    for (int32_t cell : fix) fix_col_ofs[cell / ny + 1]++;
    for (int i = 0; i < nx; ++i) fix_col_ofs[i + 1] += fix_col_ofs[i];
  }

  void destroy() {
    if (runs_dev) cudaFree(runs_dev);
    if (run_ofs_dev) cudaFree(run_ofs_dev);
    if (fix_dev) cudaFree(fix_dev);
    runs_dev = nullptr;
    run_ofs_dev = nullptr;
    fix_dev = nullptr;
  }

  // Cost-balanced tiles, one per CTA, for columns [cbeg, cend).  A CTA is done when its
  // slowest warp is; a warp pays kCol = 16 cost units per column plus mixed_cost_q units where a strip of
  // its group mixes phases (both phase bodies run), plus edge_lo_q / edge_hi_q units on the first /
  // last column of a slab whose halo column comes from a neighbour (the tile may have to wait for
  // the neighbour's flag, so it gets fewer columns).  The loops are memory-bound: measured per-CTA
  // times (STEFAN_DEV_CTA_TIMES, profiles/r2_*cta*) follow the column count, a mixed column costs
  // about a quarter column more.
  // Round 2: (1) the whole CTA budget (one wave) is handed out -- tiles go one by one to the group
  // whose tiles are currently the most expensive; (2) a group is cut into tiles of EQUAL cost.
  // (Round 1 cut greedily at a common cost T: 625 columns became 12 x 49 + 37 with 17 of 296 CTA
  // slots unused; now 14 x 44-45.)  Cached per column range.
  // all costs in SIXTEENTHS of a column (finer than the tile counts can resolve: with 21 groups and 296 CTA
  // slots a quarter column of over-estimate per mixed column took two tiles away from two plain groups)
  static constexpr int kCol = 16;
  int mixed_cost_q = 4;  // per mixed column; set by the kernel's host code before the first apply
  int min_tile_cols = 4;
  int edge_lo_q = 0, edge_hi_q = 0;
  const SegCache& segments(int cbeg, int cend, int max_ctas) {
    for (const SegCache& c : seg_cache)
      if (c.cbeg == cbeg && c.cend == cend && c.max_ctas == max_ctas) return c;
    if (const char* e = getenv("STEFAN_MIXED_COST_16")) mixed_cost_q = atoi(e);
    if (const char* e = getenv("STEFAN_MIN_TILE_COLS")) min_tile_cols = std::max(1, atoi(e));
    if (const char* e = getenv("STEFAN_EDGE_COST_16")) {
      if (edge_lo_q) edge_lo_q = atoi(e);
      if (edge_hi_q) edge_hi_q = atoi(e);
    }
    const int max_ctas_key = max_ctas;
    if (const char* e = getenv("STEFAN_CTA_BUDGET")) max_ctas = std::max(1, std::min(max_ctas, atoi(e)));  // tuning
    const int ng = ngroups, gw = group_warps, ncol = cend - cbeg;
    // cumulative cost per group: cum[g][c - cbeg] = cost of columns [cbeg, c)
    std::vector<int64_t> cum((size_t)ng * (ncol + 1), 0);
    for (int g = 0; g < ng; ++g) {
      int64_t* cg = cum.data() + (size_t)g * (ncol + 1);
      for (int c = cbeg; c < cend; ++c) {
        int mixed = 0;
        for (int sidx = g * gw; sidx < std::min(nstrips, (g + 1) * gw); ++sidx) {
          const int* mc = mixed_cum.data() + (size_t)sidx * (nx + 1);
          mixed |= mc[c + 1] - mc[c];
        }
        int64_t v = kCol + (mixed ? mixed_cost_q : 0);
        if (c == 0) v += edge_lo_q;
        if (c == nx - 1) v += edge_hi_q;
        cg[c - cbeg + 1] = cg[c - cbeg] + v;
      }
    }
    // tiles per group: hand out the budget one tile at a time to the group with the costliest tiles
    const int budget = std::max(ng, std::min(max_ctas, kMaxMarchCtas));
    const int max_per_group = std::max(1, ncol / std::max(1, min_tile_cols));
    std::vector<int> ntile(ng, 1);
    for (int total = ng; total < budget; ++total) {
      int best = -1;
      double best_cost = 0.0;
      for (int g = 0; g < ng; ++g) {
        if (ntile[g] >= max_per_group) continue;
        const double c = (double)cum[(size_t)g * (ncol + 1) + ncol] / ntile[g];
        if (c > best_cost) {
          best_cost = c;
          best = g;
        }
      }
      if (best < 0) break;
      ++ntile[best];
    }
    std::vector<int4> tiles;
    for (int g = 0; g < ng; ++g) {
      const int64_t* cg = cum.data() + (size_t)g * (ncol + 1);
      const int n = ntile[g];
      int c0 = 0;
      for (int k = 1; k <= n && c0 < ncol; ++k) {
        // end of tile k: the first column count whose cost reaches k / n of the group's cost
        int c1 = ncol;
        if (k < n) {
          const int64_t target = (cg[ncol] * k + n / 2) / n;
          c1 = (int)(std::lower_bound(cg + c0 + 1, cg + ncol + 1, target) - cg);
          // the nearer of the two columns around the target
          if (c1 > c0 + 1 && target - cg[c1 - 1] < cg[c1] - target) --c1;
          c1 = std::min(std::max(c1, c0 + 1), ncol - (n - k));  // at least one column for every tile left
          c1 = std::max(c1, c0 + 1);
        }
        tiles.push_back(make_int4(g, cbeg + c0, cbeg + c1, k - 1));
        c0 = c1;
      }
    }
    // CTA order: segment-major = CTAs with neighbouring block indices work on neighbouring
    // strips of the same columns
    std::stable_sort(tiles.begin(), tiles.end(), [](const int4& a, const int4& b) { return a.w < b.w; });
    // Experiment (STEFAN_TILE_PAIRING, default 0 = off): the two phase bodies of the order-2 / order-3 kernels together
    // exceed the instruction cache, and an SM that hosts one CTA of each phase thrashes it.  Order the tiles so that
    // the CTAs that share an SM work on the same phase: 1 = blocks b and b + n/2 are partners, 2 = blocks 2b, 2b + 1.
    static const int pairing = [] { const char* e = getenv("STEFAN_TILE_PAIRING"); return e ? atoi(e) : 0; }();
    if (pairing != 0 && tiles.size() >= 4) {
      auto phase2_share = [&](const int4& t) {
        int64_t p2 = 0, all = 0;
        for (int sidx = t.x * gw; sidx < std::min(nstrips, (t.x + 1) * gw); ++sidx) {
          const int* c2 = p2_cum.data() + (size_t)sidx * (nx + 1);
          p2 += c2[t.z] - c2[t.y];
          all += t.z - t.y;
        }
        return all ? (double)p2 / (double)all : 0.0;
      };
      std::vector<std::pair<double, int>> key(tiles.size());
      for (size_t k = 0; k < tiles.size(); ++k) key[k] = {phase2_share(tiles[k]), (int)k};
      std::stable_sort(key.begin(), key.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
      std::vector<int4> out(tiles.size());
      const size_t half = (tiles.size() + 1) / 2;
      for (size_t k = 0; k < tiles.size(); ++k) {
        const size_t dst = pairing == 1 ? (k % 2 == 0 ? k / 2 : half + k / 2) : k;
        out[std::min(dst, tiles.size() - 1)] = tiles[key[k].second];
      }
      if (pairing == 1 && tiles.size() % 2 == 1) {   // (odd count: the simple interleave above leaves a hole; fall back)
        for (size_t k = 0; k < tiles.size(); ++k) out[k] = tiles[key[k].second];
      }
      tiles.swap(out);
    }
    SegCache c;
    c.cbeg = cbeg;
    c.cend = cend;
    c.max_ctas = max_ctas_key;
    c.n = (int)tiles.size() > kMaxMarchCtas ? -1 : (int)tiles.size();
    std::memset(&c.tab, 0, sizeof(c.tab));
    for (int k = 0; k < std::max(c.n, 0); ++k) c.tab.e[k] = tiles[k];
    if (seg_cache.size() >= 8) seg_cache.erase(seg_cache.begin());
    seg_cache.push_back(c);
    return seg_cache.back();
  }
};

}  // namespace stefan
