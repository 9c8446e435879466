// oracle.cpp -- CPU restatement of the reference's normals -> RSD -> GRSD path.
// TEST INFRASTRUCTURE ONLY (see oracle.h).  Compile with -O2 -ffp-contract=off so
// that fp32 expressions are evaluated exactly as written (no FMA contraction).
//
// Reference citations are relative to /root/reference.
#include "oracle.h"

#include <algorithm>
#include <chrono>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <unordered_map>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

struct Nb {
  float d2;
  int32_t idx;
};
inline bool nb_less(const Nb& a, const Nb& b) {
  return a.d2 < b.d2 || (a.d2 == b.d2 && a.idx < b.idx);
}

inline bool finite3(const float* p) {
  return std::isfinite(p[0]) && std::isfinite(p[1]) && std::isfinite(p[2]);
}

// ---------------------------------------------------------------------------
// Uniform grid used only to make the oracle's radius search O(N k); the result
// set is defined purely by orc_d2() <= r2 and is checked against brute force.
// ---------------------------------------------------------------------------
class CellGrid {
 public:
  CellGrid(const float* xyz, int n, double r) : xyz_(xyz), n_(n) {
    double lo[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, hi[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
    int nf = 0;
    for (int i = 0; i < n; ++i) {
      const float* p = xyz + 3 * (size_t)i;
      if (!finite3(p)) continue;
      ++nf;
      for (int a = 0; a < 3; ++a) {
        lo[a] = std::min(lo[a], (double)p[a]);
        hi[a] = std::max(hi[a], (double)p[a]);
      }
    }
    if (nf == 0) {
      for (int a = 0; a < 3; ++a) lo[a] = hi[a] = 0;
    }
    cell_ = r * 1.001 + 1e-9;
    if (!(cell_ > 0)) cell_ = 1e-9;
    for (int a = 0; a < 3; ++a) {
      lo_[a] = lo[a];
      dim_[a] = (int64_t)std::floor((hi[a] - lo[a]) / cell_) + 1;
    }
    // keep the dense table bounded; otherwise use a hash map
    double cells = (double)dim_[0] * (double)dim_[1] * (double)dim_[2];
    dense_ = cells <= 6.4e7;
    std::vector<std::pair<uint64_t, int32_t>> keyed;
    keyed.reserve(nf);
    for (int i = 0; i < n; ++i) {
      const float* p = xyz + 3 * (size_t)i;
      if (!finite3(p)) continue;
      keyed.emplace_back(key_of(p), i);
    }
    std::sort(keyed.begin(), keyed.end());
    order_.resize(keyed.size());
    for (size_t i = 0; i < keyed.size(); ++i) order_[i] = keyed[i].second;
    if (dense_) {
      start_.assign((size_t)cells + 1, 0);
      for (auto& kv : keyed) start_[kv.first + 1]++;
      for (size_t c = 0; c < (size_t)cells; ++c) start_[c + 1] += start_[c];
    } else {
      size_t i = 0;
      while (i < keyed.size()) {
        size_t j = i;
        while (j < keyed.size() && keyed[j].first == keyed[i].first) ++j;
        map_[keyed[i].first] = std::make_pair((int32_t)i, (int32_t)j);
        i = j;
      }
    }
  }

  // All surface points with d2 <= r2 of q, unsorted.
  void query(const float* q, float r2, std::vector<Nb>& out) const {
    out.clear();
    if (!finite3(q)) return;
    int64_t c[3];
    for (int a = 0; a < 3; ++a) c[a] = (int64_t)std::floor(((double)q[a] - lo_[a]) / cell_);
    for (int64_t z = c[2] - 1; z <= c[2] + 1; ++z) {
      if (z < 0 || z >= dim_[2]) continue;
      for (int64_t y = c[1] - 1; y <= c[1] + 1; ++y) {
        if (y < 0 || y >= dim_[1]) continue;
        for (int64_t x = c[0] - 1; x <= c[0] + 1; ++x) {
          if (x < 0 || x >= dim_[0]) continue;
          uint64_t key = (uint64_t)((z * dim_[1] + y) * dim_[0] + x);
          int32_t b, e;
          if (dense_) {
            b = start_[key];
            e = start_[key + 1];
          } else {
            auto it = map_.find(key);
            if (it == map_.end()) continue;
            b = it->second.first;
            e = it->second.second;
          }
          for (int32_t s = b; s < e; ++s) {
            int32_t j = order_[s];
            float d2 = orc_d2(xyz_ + 3 * (size_t)j, q);
            if (d2 <= r2) out.push_back({d2, j});
          }
        }
      }
    }
  }

 private:
  uint64_t key_of(const float* p) const {
    int64_t c[3];
    for (int a = 0; a < 3; ++a) {
      c[a] = (int64_t)std::floor(((double)p[a] - lo_[a]) / cell_);
      c[a] = std::min(std::max(c[a], (int64_t)0), dim_[a] - 1);
    }
    return (uint64_t)((c[2] * dim_[1] + c[1]) * dim_[0] + c[0]);
  }
  const float* xyz_;
  int n_;
  double cell_, lo_[3];
  int64_t dim_[3];
  bool dense_;
  std::vector<int32_t> order_, start_;
  std::unordered_map<uint64_t, std::pair<int32_t, int32_t>> map_;
};

inline void sort_truncate(std::vector<Nb>& v, int max_nn) {
  std::sort(v.begin(), v.end(), nb_less);
  if (max_nn > 0 && (int)v.size() > max_nn) v.resize(max_nn);
}

inline float r2_of(double r) {
  float rf = (float)r;
  return rf * rf;
}

// ---------------------------------------------------------------------------
// 3x3 symmetric eigen-decomposition, cyclic Jacobi in double.
// a = {xx, xy, xz, yy, yz, zz}.  Eigenvalues ascending in w, eigenvectors in
// the columns of v (v[r][c]).
// ---------------------------------------------------------------------------
void eig3_jacobi(const double a[6], double w[3], double v[3][3]) {
  double m[3][3] = {{a[0], a[1], a[2]}, {a[1], a[3], a[4]}, {a[2], a[4], a[5]}};
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) v[i][j] = (i == j) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 64; ++sweep) {
    double off = std::fabs(m[0][1]) + std::fabs(m[0][2]) + std::fabs(m[1][2]);
    if (off == 0.0) break;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        if (m[p][q] == 0.0) continue;
        double theta = (m[q][q] - m[p][p]) / (2.0 * m[p][q]);
        double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        if (!std::isfinite(theta)) t = 0.0;  // |theta| overflow: rotation angle ~ 0
        double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        double apq = m[p][q];
        m[p][p] -= t * apq;
        m[q][q] += t * apq;
        m[p][q] = m[q][p] = 0.0;
        int r = 3 - p - q;
        double arp = m[r][p], arq = m[r][q];
        m[r][p] = m[p][r] = c * arp - s * arq;
        m[r][q] = m[q][r] = s * arp + c * arq;
        for (int k = 0; k < 3; ++k) {
          double vkp = v[k][p], vkq = v[k][q];
          v[k][p] = c * vkp - s * vkq;
          v[k][q] = s * vkp + c * vkq;
        }
      }
  }
  int o[3] = {0, 1, 2};
  double d[3] = {m[0][0], m[1][1], m[2][2]};
  std::sort(o, o + 3, [&](int x, int y) { return d[x] < d[y]; });
  double vv[3][3];
  for (int c = 0; c < 3; ++c) {
    w[c] = d[o[c]];
    for (int r = 0; r < 3; ++r) vv[r][c] = v[r][o[c]];
  }
  std::memcpy(v, vv, sizeof(vv));
}

// PCA normal of a neighbourhood (pcl::NormalEstimation semantics [EXTERNAL];
// call site grsd_colorCHLAC_tools.hpp:76-81).  Coordinates are taken relative to
// the query so the double sums are exact for lattice inputs.
void pca_normal(const float* xyz, const float* q, const std::vector<Nb>& nbs, const float* vp,
                float out[4], float* gap = nullptr) {
  const float nan = std::numeric_limits<float>::quiet_NaN();
  size_t k = nbs.size();
  if (k < 3) {
    out[0] = out[1] = out[2] = out[3] = nan;
    if (gap) *gap = nan;
    return;
  }
  double s1[3] = {0, 0, 0}, s2[6] = {0, 0, 0, 0, 0, 0};
  for (const Nb& nb : nbs) {
    const float* p = xyz + 3 * (size_t)nb.idx;
    double dx = (double)p[0] - (double)q[0], dy = (double)p[1] - (double)q[1],
           dz = (double)p[2] - (double)q[2];
    s1[0] += dx;
    s1[1] += dy;
    s1[2] += dz;
    s2[0] += dx * dx;
    s2[1] += dx * dy;
    s2[2] += dx * dz;
    s2[3] += dy * dy;
    s2[4] += dy * dz;
    s2[5] += dz * dz;
  }
  double inv = 1.0 / (double)k;
  double mx = s1[0] * inv, my = s1[1] * inv, mz = s1[2] * inv;
  double cov[6] = {s2[0] * inv - mx * mx, s2[1] * inv - mx * my, s2[2] * inv - mx * mz,
                   s2[3] * inv - my * my, s2[4] * inv - my * mz, s2[5] * inv - mz * mz};
  double w[3], v[3][3];
  eig3_jacobi(cov, w, v);
  double nx = v[0][0], ny = v[1][0], nz = v[2][0];
  double len = std::sqrt(nx * nx + ny * ny + nz * nz);
  nx /= len;
  ny /= len;
  nz /= len;
  // flipNormalTowardsViewpoint: n.(vp - p) >= 0
  double dot = nx * ((double)vp[0] - q[0]) + ny * ((double)vp[1] - q[1]) + nz * ((double)vp[2] - q[2]);
  if (dot < 0) {
    nx = -nx;
    ny = -ny;
    nz = -nz;
  }
  double tr = w[0] + w[1] + w[2];
  // conditioning of the normal: a perturbation e of the covariance turns it by about e / (l1 - l0)
  if (gap) *gap = (tr != 0.0) ? (float)((w[1] - w[0]) / tr) : 0.0f;
  out[0] = (float)nx;
  out[1] = (float)ny;
  out[2] = (float)nz;
  out[3] = (tr != 0.0) ? (float)std::fabs(w[0] / tr) : 0.0f;
}

// fp32 dot product evaluated as the reference writes it
// (radius_estimation.cpp:153-155: float*float + float*float + float*float,
// left to right, then widened to double).
inline double cosine_f32(const float* a, const float* b) {
  float c = a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
  return (double)c;
}

// ---------------------------------------------------------------------------
// RSD core.  In-tree variant: radius_estimation.cpp:140-202.  The neighbour list
// is sorted by (d2, idx); `self` is excluded by index (the
// reference skips list position 0, which is the query itself for distinct
// points).  REF_IS_NEAREST restates pcl::RSDEstimation [EXTERNAL]: the reference
// element is the first (nearest) neighbour, distances are measured from it and
// neighbours farther than the radius from it are skipped.
// ---------------------------------------------------------------------------
void rsd_core(const float* xyz, const float* nrm, int nstride, const std::vector<Nb>& nbs,
              int self, const float* self_normal, double radius, int ndiv, double plane_radius,
              int flags, double* min_radius_out, double* max_radius_out) {
  std::vector<double> mn(ndiv, +DBL_MAX), mx(ndiv, -DBL_MAX);
  if ((flags & ORC_RSD_SEED_BIN0) && ndiv > 0) mn[0] = mx[0] = 0.0;
  const bool nearest = (flags & ORC_RSD_REF_IS_NEAREST) != 0;
  const float* ref_n = self_normal;
  const float* ref_p = nullptr;
  size_t first = 0;
  if (nearest) {
    if (nbs.empty()) {
      *min_radius_out = *max_radius_out = plane_radius;
      return;
    }
    ref_n = nrm + (size_t)nstride * nbs[0].idx;
    ref_p = xyz + 3 * (size_t)nbs[0].idx;
    first = 1;
  }
  for (size_t ni = first; ni < nbs.size(); ++ni) {
    int j = nbs[ni].idx;
    if (!nearest && j == self) continue;
    double cosine = cosine_f32(ref_n, nrm + (size_t)nstride * j);
    if (cosine > 1) cosine = 1;  // radius_estimation.cpp:158-159
    if (cosine < -1) cosine = -1;
    double angle = std::acos(cosine);
    if (angle > M_PI / 2) angle = M_PI - angle;  // :161
    double dist;
    if (nearest) {
      dist = std::sqrt((double)orc_d2(xyz + 3 * (size_t)j, ref_p));
      if (dist > radius) continue;
    } else {
      // :165 `sqrt (points_sqr_distances_[cp][ni])` on a float under `using namespace std;` (:4) is std::sqrt(float)
      dist = (double)std::sqrt(nbs[ni].d2);
    }
    int bin = (int)std::floor(ndiv * dist / radius);  // :168
    if (bin > ndiv - 1) bin = ndiv - 1;  // epsilon rule: the reference indexes out of bounds here
    if (mn[bin] > angle) mn[bin] = angle;  // :171-172 (NaN falls through both)
    if (mx[bin] < angle) mx[bin] = angle;
  }
  double Amint_Amin = 0, Amint_d = 0, Amaxt_Amax = 0, Amaxt_d = 0;
  for (int di = 0; di < ndiv; ++di) {
    if (mx[di] >= 0) {  // :181
      double p_min = mn[di], p_max = mx[di];
      double f = (di + 0.5) * radius / ndiv;
      Amint_Amin += p_min * p_min;
      Amint_d += p_min * f;
      Amaxt_Amax += p_max * p_max;
      Amaxt_d += p_max * f;
    }
  }
  double max_radius = (Amint_Amin == 0) ? plane_radius : std::min(Amint_d / Amint_Amin, plane_radius);
  double min_radius = (Amaxt_Amax == 0) ? plane_radius : std::min(Amaxt_d / Amaxt_Amax, plane_radius);
  if (flags & ORC_RSD_SCALE_SORT) {
    // newer pcl::computeRSD [EXTERNAL]: computed in float, scaled, then ordered
    float a = (float)max_radius, b = (float)min_radius;  // a: from min-angle line, b: from max-angle line
    a *= 1.1f;
    b *= 0.9f;
    if (a < b) {
      min_radius = a;
      max_radius = b;
    } else {
      min_radius = b;
      max_radius = a;
    }
  }
  *min_radius_out = min_radius;
  *max_radius_out = max_radius;
}

// ---------------------------------------------------------------------------
// kd-tree for the reference-faithful timing mode (stands in for
// cloud_kdtree::KdTreeANN, radius_estimation.cpp:107).  Median split on the
// widest axis, small leaf buckets, iterative radius search.
// ---------------------------------------------------------------------------
class KdTree {
 public:
  KdTree(const float* xyz, int n) : xyz_(xyz) {
    idx_.reserve(n);
    for (int i = 0; i < n; ++i)
      if (finite3(xyz + 3 * (size_t)i)) idx_.push_back(i);
    nodes_.reserve(idx_.size() / 4 + 16);
    if (!idx_.empty()) build(0, (int)idx_.size());
  }
  // The k smallest (d2, index) pairs over all points, ascending (cloud_kdtree's nearestKSearch as used at
  // noise_removal.cpp:90 [EXTERNAL semantics]; ties by input index).
  void knn(const float* q, int k, std::vector<Nb>& out) const {
    out.clear();
    if (nodes_.empty() || !finite3(q) || k <= 0) return;
    knn_rec(0, q, k, out);
    std::sort_heap(out.begin(), out.end(), nb_less);
  }
  void radius(const float* q, float r, float r2, std::vector<Nb>& out) const {
    out.clear();
    if (nodes_.empty() || !finite3(q)) return;
    int stack[128], sp = 0;
    stack[sp++] = 0;
    while (sp) {
      const Node& nd = nodes_[stack[--sp]];
      if (nd.axis < 0) {
        for (int s = nd.lo; s < nd.hi; ++s) {
          int j = idx_[s];
          float d2 = orc_d2(xyz_ + 3 * (size_t)j, q);
          if (d2 <= r2) out.push_back({d2, j});
        }
        continue;
      }
      float diff = q[nd.axis] - nd.split;
      // conservative pruning (slack keeps boundary points reachable)
      float slack = r * 1.0001f + 1e-12f;
      if (diff <= slack) stack[sp++] = nd.left;
      if (diff >= -slack) stack[sp++] = nd.right;
    }
  }

 private:
  struct Node {
    int axis;  // -1 leaf
    float split;
    int left, right, lo, hi;
  };
  void knn_rec(int node, const float* q, int k, std::vector<Nb>& heap) const {  // max-heap under nb_less
    const Node& nd = nodes_[node];
    if (nd.axis < 0) {
      for (int s = nd.lo; s < nd.hi; ++s) {
        const int j = idx_[s];
        const Nb c{orc_d2(xyz_ + 3 * (size_t)j, q), j};
        if ((int)heap.size() < k) {
          heap.push_back(c);
          std::push_heap(heap.begin(), heap.end(), nb_less);
        } else if (nb_less(c, heap.front())) {
          std::pop_heap(heap.begin(), heap.end(), nb_less);
          heap.back() = c;
          std::push_heap(heap.begin(), heap.end(), nb_less);
        }
      }
      return;
    }
    const double diff = (double)q[nd.axis] - (double)nd.split;
    const int near = diff < 0 ? nd.left : nd.right, far = diff < 0 ? nd.right : nd.left;
    knn_rec(near, q, k, heap);
    // conservative pruning: the far side can only matter if the plane is not farther than the current worst
    if ((int)heap.size() < k || diff * diff <= (double)heap.front().d2 * 1.0001 + 1e-30) knn_rec(far, q, k, heap);
  }
  int build(int lo, int hi) {
    int id = (int)nodes_.size();
    nodes_.push_back(Node{-1, 0.f, -1, -1, lo, hi});
    if (hi - lo <= 8) return id;
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int s = lo; s < hi; ++s) {
      const float* p = xyz_ + 3 * (size_t)idx_[s];
      for (int a = 0; a < 3; ++a) {
        mn[a] = std::min(mn[a], p[a]);
        mx[a] = std::max(mx[a], p[a]);
      }
    }
    int axis = 0;
    if (mx[1] - mn[1] > mx[axis] - mn[axis]) axis = 1;
    if (mx[2] - mn[2] > mx[axis] - mn[axis]) axis = 2;
    if (!(mx[axis] > mn[axis])) return id;  // all duplicates: keep as leaf
    int mid = (lo + hi) / 2;
    std::nth_element(idx_.begin() + lo, idx_.begin() + mid, idx_.begin() + hi, [&](int a, int b) {
      return xyz_[3 * (size_t)a + axis] < xyz_[3 * (size_t)b + axis];
    });
    float split = xyz_[3 * (size_t)idx_[mid] + axis];
    int l = build(lo, mid);
    int r = build(mid, hi);
    nodes_[id].axis = axis;
    nodes_[id].split = split;
    nodes_[id].left = l;
    nodes_[id].right = r;
    return id;
  }
  const float* xyz_;
  std::vector<int> idx_;
  std::vector<Node> nodes_;
};

int resolve_threads(int nthreads) {
#ifdef _OPENMP
  if (nthreads <= 0) nthreads = omp_get_max_threads();
  return nthreads;
#else
  (void)nthreads;
  return 1;
#endif
}

// 13 "half" offsets then their negations, grsd_colorCHLAC_tools.hpp:187-222
void fill_offsets26(int32_t off[26][3]) {
  int idx = 0;
  for (int i = -1; i < 2; i++)
    for (int j = -1; j < 2; j++) {
      off[idx][0] = i;
      off[idx][1] = j;
      off[idx][2] = -1;
      idx++;
    }
  for (int i = -1; i < 2; i++) {
    off[idx][0] = i;
    off[idx][1] = -1;
    off[idx][2] = 0;
    idx++;
  }
  off[idx][0] = -1;
  off[idx][1] = 0;
  off[idx][2] = 0;
  for (int c = 0; c < 13; ++c)
    for (int a = 0; a < 3; ++a) off[13 + c][a] = -off[c][a];
}

// pcl::VoxelGrid::getNeighborCentroidIndices [EXTERNAL], call sites grsd_colorCHLAC_tools.hpp:249 and
// color_chlac.hpp:1512.  ijk = floor(ref / leaf): a division in the PCL of the reference's era -- the shipped cube and
// dice feature vectors, whose voxel centroids sit on voxel faces, give 2187 / 2809 occupied half-stencil pairs with it
// and 2157 / 2789 with ref * (1 / leaf) (tests/test_color_chlac.py).
inline int neighbor_centroid(const float* ref, float leaf, const int32_t* min_b,
                             const int32_t* div_b, const int32_t* layout, const int32_t* disp) {
  int ijk[3], max_b[3];
  for (int a = 0; a < 3; ++a) {
    ijk[a] = (int)std::floor(ref[a] / leaf);
    max_b[a] = min_b[a] + div_b[a] - 1;
  }
  for (int a = 0; a < 3; ++a) {
    if (!(min_b[a] - ijk[a] <= disp[a] && max_b[a] - ijk[a] >= disp[a])) return -1;
  }
  int64_t lin = (int64_t)(ijk[0] + disp[0] - min_b[0]) +
                (int64_t)(ijk[1] + disp[1] - min_b[1]) * div_b[0] +
                (int64_t)(ijk[2] + disp[2] - min_b[2]) * div_b[0] * div_b[1];
  return layout[lin];
}

}  // namespace

extern "C" {

float orc_d2(const float* a, const float* b) {
  float dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
  float xx = dx * dx, yy = dy * dy, zz = dz * dz;
  float s = xx + yy;
  return s + zz;
}

int orc_num_threads(void) { return resolve_threads(0); }

int64_t orc_radius_search(const float* surface_xyz, int n, const float* query_xyz, int nq,
                          double r, int max_nn, int64_t* offsets, int32_t* idx, float* d2,
                          int64_t cap, int nthreads) {
  CellGrid grid(surface_xyz, n, r);
  const float r2 = r2_of(r);
  nthreads = resolve_threads(nthreads);
  std::vector<int64_t> counts(nq, 0);
#pragma omp parallel num_threads(nthreads)
  {
    std::vector<Nb> nbs;
#pragma omp for schedule(dynamic, 512)
    for (int qi = 0; qi < nq; ++qi) {
      grid.query(query_xyz + 3 * (size_t)qi, r2, nbs);
      int64_t k = (int64_t)nbs.size();
      if (max_nn > 0 && k > max_nn) k = max_nn;
      counts[qi] = k;
    }
  }
  offsets[0] = 0;
  for (int qi = 0; qi < nq; ++qi) offsets[qi + 1] = offsets[qi] + counts[qi];
  int64_t total = offsets[nq];
  if (total > cap || (idx == nullptr && d2 == nullptr)) return total;
#pragma omp parallel num_threads(nthreads)
  {
    std::vector<Nb> nbs;
#pragma omp for schedule(dynamic, 512)
    for (int qi = 0; qi < nq; ++qi) {
      grid.query(query_xyz + 3 * (size_t)qi, r2, nbs);
      sort_truncate(nbs, max_nn);
      int64_t o = offsets[qi];
      for (size_t s = 0; s < nbs.size(); ++s) {
        if (idx) idx[o + s] = nbs[s].idx;
        if (d2) d2[o + s] = nbs[s].d2;
      }
    }
  }
  return total;
}

int64_t orc_radius_search_brute(const float* surface_xyz, int n, const float* query_xyz, int nq,
                                double r, int max_nn, int64_t* offsets, int32_t* idx, float* d2,
                                int64_t cap) {
  const float r2 = r2_of(r);
  std::vector<Nb> nbs;
  int64_t total = 0;
  offsets[0] = 0;
  for (int qi = 0; qi < nq; ++qi) {
    nbs.clear();
    const float* q = query_xyz + 3 * (size_t)qi;
    for (int j = 0; j < n; ++j) {
      float dd = orc_d2(surface_xyz + 3 * (size_t)j, q);
      if (dd <= r2) nbs.push_back({dd, j});
    }
    sort_truncate(nbs, max_nn);
    for (size_t s = 0; s < nbs.size(); ++s) {
      if (total + (int64_t)s < cap) {
        if (idx) idx[total + s] = nbs[s].idx;
        if (d2) d2[total + s] = nbs[s].d2;
      }
    }
    total += (int64_t)nbs.size();
    offsets[qi + 1] = total;
  }
  return total;
}

int orc_normals(const float* xyz, int n, double r, int max_nn, const float* vp, float* out_n4,
                int32_t* out_k, int nthreads) {
  CellGrid grid(xyz, n, r);
  const float r2 = r2_of(r);
  const float zero[3] = {0, 0, 0};
  if (!vp) vp = zero;
  nthreads = resolve_threads(nthreads);
#pragma omp parallel num_threads(nthreads)
  {
    std::vector<Nb> nbs;
#pragma omp for schedule(dynamic, 512)
    for (int i = 0; i < n; ++i) {
      const float* q = xyz + 3 * (size_t)i;
      grid.query(q, r2, nbs);
      if (max_nn > 0 && (int)nbs.size() > max_nn) sort_truncate(nbs, max_nn);
      pca_normal(xyz, q, nbs, vp, out_n4 + 4 * (size_t)i);
      if (out_k) out_k[i] = (int32_t)nbs.size();
    }
  }
  return 0;
}

// The normals of orc_normals plus their conditioning (l1 - l0) / (l0 + l1 + l2): the parity gates list the points whose
// own decision margin is below the noise of any fp32 implementation instead of dropping them.
int orc_normals_gap(const float* xyz, int n, double r, int max_nn, const float* vp, float* out_n4, float* out_gap,
                    int nthreads) {
  CellGrid grid(xyz, n, r);
  const float r2 = r2_of(r);
  const float zero[3] = {0, 0, 0};
  if (!vp) vp = zero;
  nthreads = resolve_threads(nthreads);
#pragma omp parallel num_threads(nthreads)
  {
    std::vector<Nb> nbs;
#pragma omp for schedule(dynamic, 512)
    for (int i = 0; i < n; ++i) {
      const float* q = xyz + 3 * (size_t)i;
      grid.query(q, r2, nbs);
      if (max_nn > 0 && (int)nbs.size() > max_nn) sort_truncate(nbs, max_nn);
      pca_normal(xyz, q, nbs, vp, out_n4 + 4 * (size_t)i, out_gap + i);
    }
  }
  return 0;
}

// k-NN normals: nearestKSearch (i, k_) + computePointNormal + flipNormalTowardsViewpoint as in
// cloud_tools/src/table_object_detector_passive.cpp:668-714 and cloud_algos/src/cylinder_fit_algo.cpp:138-203
// [point_cloud_mapping, EXTERNAL]: the PCA of orc_normals over the k smallest (d2, index) pairs, the query included.
int orc_normals_knn(const float* xyz, int n, int k, const float* vp, float* out_n4, int nthreads) {
  if (k < 3) return -1;
  KdTree tree(xyz, n);
  int nfinite = 0;
  for (int i = 0; i < n; ++i) nfinite += finite3(xyz + 3 * (size_t)i) ? 1 : 0;
  if (k > nfinite) return -2;
  const float zero[3] = {0, 0, 0};
  if (!vp) vp = zero;
  nthreads = resolve_threads(nthreads);
#pragma omp parallel num_threads(nthreads)
  {
    std::vector<Nb> nbs;
#pragma omp for schedule(dynamic, 256)
    for (int i = 0; i < n; ++i) {
      const float* q = xyz + 3 * (size_t)i;
      tree.knn(q, k, nbs);  // empty for a non-finite query: pca_normal then writes NaN
      pca_normal(xyz, q, nbs, vp, out_n4 + 4 * (size_t)i);
    }
  }
  return 0;
}

int orc_rsd(const float* xyz, const float* normals, int normal_stride, int n, double r,
            int max_nn, int ndiv, double plane_radius, int flags, float* r_min, float* r_max,
            float* r_dif, int nthreads) {
  if (ndiv <= 0) return -1;
  CellGrid grid(xyz, n, r);
  const float r2 = r2_of(r);
  nthreads = resolve_threads(nthreads);
#pragma omp parallel num_threads(nthreads)
  {
    std::vector<Nb> nbs;
#pragma omp for schedule(dynamic, 512)
    for (int i = 0; i < n; ++i) {
      grid.query(xyz + 3 * (size_t)i, r2, nbs);
      sort_truncate(nbs, max_nn);
      double mn, mx;
      rsd_core(xyz, normals, normal_stride, nbs, i, normals + (size_t)normal_stride * i, r, ndiv,
               plane_radius, flags, &mn, &mx);
      if (r_min) r_min[i] = (float)mn;  // radius_estimation.cpp:200-202
      if (r_max) r_max[i] = (float)mx;
      if (r_dif) r_dif[i] = (float)(mx - mn);
    }
  }
  return 0;
}

int orc_rsd_queries(const float* surface_xyz, const float* normals, int normal_stride, int n,
                    const float* query_xyz, int nq, double r, int max_nn, int ndiv,
                    double plane_radius, int flags, float* r_min, float* r_max, int nthreads) {
  if (ndiv <= 0) return -1;
  CellGrid grid(surface_xyz, n, r);
  const float r2 = r2_of(r);
  nthreads = resolve_threads(nthreads);
  const float zero_n[3] = {0, 0, 0};
#pragma omp parallel num_threads(nthreads)
  {
    std::vector<Nb> nbs;
#pragma omp for schedule(dynamic, 64)
    for (int i = 0; i < nq; ++i) {
      grid.query(query_xyz + 3 * (size_t)i, r2, nbs);
      sort_truncate(nbs, max_nn);
      double mn, mx;
      rsd_core(surface_xyz, normals, normal_stride, nbs, -1, zero_n, r, ndiv, plane_radius,
               flags | ORC_RSD_REF_IS_NEAREST, &mn, &mx);
      r_min[i] = (float)mn;
      r_max[i] = (float)mx;
    }
  }
  return 0;
}

int orc_rsd_ref_faithful(const float* xyz, const float* normals, int normal_stride, int n,
                         double r, int max_nn, int ndiv, double plane_radius, float* r_min,
                         float* r_max, double* phase_s) {
  using clk = std::chrono::steady_clock;
  auto secs = [](clk::time_point a, clk::time_point b) {
    return std::chrono::duration<double>(b - a).count();
  };
  const float r2 = r2_of(r);
  auto t0 = clk::now();
  KdTree tree(xyz, n);  // radius_estimation.cpp:103-109
  auto t1 = clk::now();
  // materialised neighbour lists, radius_estimation.cpp:112-124
  std::vector<std::vector<int>> points_indices(n);
  std::vector<std::vector<float>> points_sqr_distances(n);
  std::vector<Nb> nbs;
  for (int cp = 0; cp < n; ++cp) {
    tree.radius(xyz + 3 * (size_t)cp, (float)r, r2, nbs);
    sort_truncate(nbs, max_nn);
    points_indices[cp].resize(nbs.size());
    points_sqr_distances[cp].resize(nbs.size());
    for (size_t s = 0; s < nbs.size(); ++s) {
      points_indices[cp][s] = nbs[s].idx;
      points_sqr_distances[cp][s] = nbs[s].d2;
    }
  }
  auto t2 = clk::now();
  for (int cp = 0; cp < n; ++cp) {  // radius_estimation.cpp:140-215
    nbs.resize(points_indices[cp].size());
    for (size_t s = 0; s < nbs.size(); ++s) nbs[s] = {points_sqr_distances[cp][s], points_indices[cp][s]};
    double mn, mx;
    rsd_core(xyz, normals, normal_stride, nbs, cp, normals + (size_t)normal_stride * cp, r, ndiv,
             plane_radius, 0, &mn, &mx);
    if (r_min) r_min[cp] = (float)mn;
    if (r_max) r_max[cp] = (float)mx;
  }
  auto t3 = clk::now();
  if (phase_s) {
    phase_s[0] = secs(t0, t1);
    phase_s[1] = secs(t1, t2);
    phase_s[2] = secs(t2, t3);
  }
  return 0;
}

int orc_voxel_grid(const float* xyz, int n, float leaf, int32_t* min_b, int32_t* div_b,
                   float* centroids, int32_t* layout, int32_t* counts) {
  // pcl::VoxelGrid::applyFilter [EXTERNAL]: inverse leaf in fp32, fp32 multiply, floor.
  const float inv = 1.0f / leaf;
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  int nf = 0;
  for (int i = 0; i < n; ++i) {
    const float* p = xyz + 3 * (size_t)i;
    if (!finite3(p)) continue;
    ++nf;
    for (int a = 0; a < 3; ++a) {
      mn[a] = std::min(mn[a], p[a]);
      mx[a] = std::max(mx[a], p[a]);
    }
  }
  if (nf == 0) {
    for (int a = 0; a < 3; ++a) min_b[a] = div_b[a] = 0;
    return 0;
  }
  for (int a = 0; a < 3; ++a) {
    min_b[a] = (int32_t)std::floor(mn[a] * inv);
    int32_t max_b = (int32_t)std::floor(mx[a] * inv);
    div_b[a] = max_b - min_b[a] + 1;
  }
  int64_t cells = (int64_t)div_b[0] * div_b[1] * div_b[2];
  std::vector<std::pair<int64_t, int32_t>> keyed;
  keyed.reserve(nf);
  for (int i = 0; i < n; ++i) {
    const float* p = xyz + 3 * (size_t)i;
    if (!finite3(p)) continue;
    int ijk0 = (int)(std::floor(p[0] * inv) - min_b[0]);
    int ijk1 = (int)(std::floor(p[1] * inv) - min_b[1]);
    int ijk2 = (int)(std::floor(p[2] * inv) - min_b[2]);
    keyed.emplace_back((int64_t)ijk0 + (int64_t)ijk1 * div_b[0] + (int64_t)ijk2 * div_b[0] * div_b[1], i);
  }
  std::sort(keyed.begin(), keyed.end());
  int nvox = 0;
  if (layout)
    for (int64_t c = 0; c < cells; ++c) layout[c] = -1;
  size_t i = 0;
  while (i < keyed.size()) {
    size_t j = i;
    // Eigen::VectorXf centroid: float sums in cloud order; `centroid /= (float)nr_points` multiplies by the
    // reciprocal in the Eigen 3.0-3.2 the reference was built with (DenseBase::operator/= for non-integer scalars).
    // The reference's shipped feature vectors show it: a voxel of k points of colour 255 comes out as 254 for some k,
    // and the cube's centroids on voxel faces fall into the neighbouring voxel.
    float s[3] = {0.f, 0.f, 0.f};
    while (j < keyed.size() && keyed[j].first == keyed[i].first) {
      const float* p = xyz + 3 * (size_t)keyed[j].second;
      s[0] += p[0];
      s[1] += p[1];
      s[2] += p[2];
      ++j;
    }
    if (centroids) {
      const float rcp = 1.0f / (float)(j - i);
      for (int a = 0; a < 3; ++a) centroids[3 * (size_t)nvox + a] = s[a] * rcp;
    }
    if (layout) layout[keyed[i].first] = nvox;
    if (counts) counts[nvox] = (int32_t)(j - i);
    ++nvox;
    i = j;
  }
  return nvox;
}

int orc_get_type(float min_radius, float max_radius) {
  // grsd_colorCHLAC_tools.hpp:104-116; class ids grsd_colorCHLAC_tools.h:10-16
  if (min_radius > 0.100)
    return 1;  // PLANE
  else if (max_radius > 0.175)
    return 2;  // CYLINDER
  else if (min_radius < 0.015)
    return 0;  // NOISE
  else if (max_radius - min_radius < 0.050)
    return 3;  // SPHERE
  else
    return 4;  // EDGE
}

void orc_offsets26(int32_t* out) {
  int32_t off[26][3];
  fill_offsets26(off);
  std::memcpy(out, off, sizeof(off));
}

int orc_grsd_transitions(const float* centroids, int nvox, const int32_t* types, float leaf,
                         const int32_t* min_b, const int32_t* div_b, const int32_t* layout,
                         int32_t* transition36, int32_t* hist21) {
  int32_t off[26][3];
  fill_offsets26(off);
  int32_t M[6][6];
  std::memset(M, 0, sizeof(M));
  for (int v = 0; v < nvox; ++v) {  // grsd_colorCHLAC_tools.hpp:230-260, hist_num == 1
    int src = types[v];
    for (int o = 0; o < 26; ++o) {
      int nb = neighbor_centroid(centroids + 3 * (size_t)v, leaf, min_b, div_b, layout, off[o]);
      int nt = (nb == -1) ? 5 : types[nb];
      M[src][nt]++;
    }
  }
  if (transition36) std::memcpy(transition36, M, sizeof(M));
  if (hist21) {  // :266-276
    int nrf = 0;
    for (int i = 0; i < 6; ++i)
      for (int j = i; j < 6; ++j) hist21[nrf++] = M[i][j];
  }
  return 0;
}

int orc_grsd21_subdiv(const float* centroids, int nvox, const int32_t* types, float leaf,
                      const int32_t* min_b, const int32_t* div_b, const int32_t* layout,
                      int subdivision_size, int off_x, int off_y, int off_z, int32_t* subdiv_b,
                      int32_t* hist21) {
  // grsd_colorCHLAC_tools.hpp:140-161
  if (subdivision_size < 0) return -1;
  int hist_num = 1;
  float inverse_subdivision_size = 0.f;
  int sb[3] = {1, 1, 1};
  if (subdivision_size > 0) {
    inverse_subdivision_size = 1.0 / subdivision_size;
    if (div_b[0] <= off_x || div_b[1] <= off_y || div_b[2] <= off_z) {
      if (subdiv_b) subdiv_b[0] = subdiv_b[1] = subdiv_b[2] = 0;
      return 0;
    }
    sb[0] = (int)std::ceil((div_b[0] - off_x) * inverse_subdivision_size);
    sb[1] = (int)std::ceil((div_b[1] - off_y) * inverse_subdivision_size);
    sb[2] = (int)std::ceil((div_b[2] - off_z) * inverse_subdivision_size);
    hist_num = sb[0] * sb[1] * sb[2];
  }
  if (subdiv_b) {
    subdiv_b[0] = sb[0];
    subdiv_b[1] = sb[1];
    subdiv_b[2] = sb[2];
  }
  if (!hist21) return hist_num;
  int32_t off[26][3];
  fill_offsets26(off);
  std::vector<int32_t> M((size_t)hist_num * 36, 0);
  for (int v = 0; v < nvox; ++v) {
    int hist_idx = 0;
    const float* c = centroids + 3 * (size_t)v;
    if (hist_num != 1) {  // :233-246
      const int tmp_x = std::floor(c[0] / leaf) - min_b[0] - off_x;
      const int tmp_y = std::floor(c[1] / leaf) - min_b[1] - off_y;
      const int tmp_z = std::floor(c[2] / leaf) - min_b[2] - off_z;
      if (tmp_x < 0 || tmp_y < 0 || tmp_z < 0) continue;
      int ix = (int)std::floor(tmp_x * inverse_subdivision_size);
      int iy = (int)std::floor(tmp_y * inverse_subdivision_size);
      int iz = (int)std::floor(tmp_z * inverse_subdivision_size);
      if (ix >= sb[0] || iy >= sb[1] || iz >= sb[2]) continue;  // the reference indexes past its vector here (UB): voxel left out
      hist_idx = ix + iy * sb[0] + iz * sb[0] * sb[1];
    }
    int src = types[v];
    for (int o = 0; o < 26; ++o) {
      int nb = neighbor_centroid(c, leaf, min_b, div_b, layout, off[o]);
      int nt = (nb == -1) ? 5 : types[nb];
      M[(size_t)hist_idx * 36 + src * 6 + nt]++;
    }
  }
  for (int h = 0; h < hist_num; ++h) {
    int nrf = 0;
    for (int i = 0; i < 6; ++i)
      for (int j = i; j < 6; ++j) hist21[(size_t)h * 21 + nrf++] = M[(size_t)h * 36 + i * 6 + j];
  }
  return hist_num;
}

int orc_grsd21(const float* xyz, const float* normals_in, int normal_stride, int n, float leaf,
               double r_normals, double rsd_radius_min, int rsd_flags, const float* vp,
               int32_t* hist21, int32_t* labels, float* radii, int32_t cap_vox,
               int32_t* nvox_out, int nthreads) {
  std::vector<float> n4;
  const float* nrm = normals_in;
  int stride = normal_stride;
  if (!nrm) {
    n4.resize((size_t)n * 4);
    orc_normals(xyz, n, r_normals, 0, vp, n4.data(), nullptr, nthreads);
    nrm = n4.data();
    stride = 4;
  }
  int32_t min_b[3], div_b[3];
  int nvox = orc_voxel_grid(xyz, n, leaf, min_b, div_b, nullptr, nullptr, nullptr);
  if (nvox_out) *nvox_out = nvox;
  std::vector<float> cent((size_t)nvox * 3);
  std::vector<int32_t> layout((size_t)div_b[0] * div_b[1] * div_b[2]);
  orc_voxel_grid(xyz, n, leaf, min_b, div_b, cent.data(), layout.data(), nullptr);
  // grsd_colorCHLAC_tools.hpp:172: std::max(rsd_radius_search, voxel_size/2 * sqrt(3)),
  // float/2 * double -> double
  double r_rsd = std::max(rsd_radius_min, leaf / 2 * std::sqrt(3));
  std::vector<float> rmin(nvox), rmax(nvox);
  // PCL RSDEstimation defaults, never overridden in the tree: nr_subdiv 5, plane_radius 0.2
  orc_rsd_queries(xyz, nrm, stride, n, cent.data(), nvox, r_rsd, 0, 5, 0.2, rsd_flags, rmin.data(),
                  rmax.data(), nthreads);
  std::vector<int32_t> types(nvox);
  for (int v = 0; v < nvox; ++v) types[v] = orc_get_type(rmin[v], rmax[v]);  // :225-228
  for (int v = 0; v < nvox && v < cap_vox; ++v) {
    if (labels) labels[v] = types[v];
    if (radii) {
      radii[2 * v] = rmin[v];
      radii[2 * v + 1] = rmax[v];
    }
  }
  return orc_grsd_transitions(cent.data(), nvox, types.data(), leaf, min_b, div_b, layout.data(),
                              nullptr, hist21);
}

int orc_voxel_normals(const float* xyz, const float* normals, int normal_stride, int n, float leaf,
                      float* out) {
  // same keys and voxel order as orc_voxel_grid
  const float inv = 1.0f / leaf;
  int32_t min_b[3], div_b[3];
  int nvox = orc_voxel_grid(xyz, n, leaf, min_b, div_b, nullptr, nullptr, nullptr);
  if (nvox == 0) return 0;
  std::vector<std::pair<int64_t, int32_t>> keyed;
  for (int i = 0; i < n; ++i) {
    const float* p = xyz + 3 * (size_t)i;
    if (!finite3(p)) continue;
    int ijk0 = (int)(std::floor(p[0] * inv) - min_b[0]);
    int ijk1 = (int)(std::floor(p[1] * inv) - min_b[1]);
    int ijk2 = (int)(std::floor(p[2] * inv) - min_b[2]);
    keyed.emplace_back((int64_t)ijk0 + (int64_t)ijk1 * div_b[0] + (int64_t)ijk2 * div_b[0] * div_b[1], i);
  }
  std::sort(keyed.begin(), keyed.end());
  int v = 0;
  size_t i = 0;
  while (i < keyed.size()) {
    size_t j = i;
    float s[3] = {0.f, 0.f, 0.f};  // part of the same Eigen::VectorXf as the centroid (orc_voxel_grid)
    while (j < keyed.size() && keyed[j].first == keyed[i].first) {
      const float* q = normals + (size_t)normal_stride * keyed[j].second;
      s[0] += q[0];
      s[1] += q[1];
      s[2] += q[2];
      ++j;
    }
    const float rcp = 1.0f / (float)(j - i);
    for (int a = 0; a < 3; ++a) out[3 * (size_t)v + a] = s[a] * rcp;
    ++v;
    i = j;
  }
  return v;
}

int orc_grsd_signature(int kind, const float* centroids, const float* cent_normals, int nvox,
                       const int32_t* types, float leaf, const int32_t* min_b, const int32_t* div_b,
                       const int32_t* layout, int subdivision_size, int off_x, int off_y, int off_z,
                       int32_t* subdiv_b, int32_t* hist) {
  if (kind < 0 || kind > 2) return -2;
  if (kind == ORC_SIG_PLUSGRSD110 && !cent_normals && hist) return -2;
  // subdivision bookkeeping, identical in the three extractors (:140-161, :322-343, :479-500)
  if (subdivision_size < 0) return -1;
  int hist_num = 1;
  float inverse_subdivision_size = 0.f;
  int sb[3] = {1, 1, 1};
  if (subdivision_size > 0) {
    inverse_subdivision_size = 1.0 / subdivision_size;
    if (div_b[0] <= off_x || div_b[1] <= off_y || div_b[2] <= off_z) {
      if (subdiv_b) subdiv_b[0] = subdiv_b[1] = subdiv_b[2] = 0;
      return 0;
    }
    sb[0] = (int)std::ceil((div_b[0] - off_x) * inverse_subdivision_size);
    sb[1] = (int)std::ceil((div_b[1] - off_y) * inverse_subdivision_size);
    sb[2] = (int)std::ceil((div_b[2] - off_z) * inverse_subdivision_size);
    hist_num = sb[0] * sb[1] * sb[2];
  }
  if (subdiv_b) {
    subdiv_b[0] = sb[0];
    subdiv_b[1] = sb[1];
    subdiv_b[2] = sb[2];
  }
  if (!hist) return hist_num;
  const int dim = kind == ORC_SIG_GRSD21 ? 21 : (kind == ORC_SIG_GRSD325 ? 325 : 110);
  int32_t off[26][3];
  fill_offsets26(off);
  const int NRDIV = 7, NRCLASS = 5;  // grsd_colorCHLAC_tools.h:9,18
  // PlusGRSD: centroid normals re-normalised in fp32 (:559-560), Eigen normalize() = v / sqrt(v.v)
  std::vector<float> nn;
  if (kind == ORC_SIG_PLUSGRSD110) {
    nn.resize((size_t)nvox * 3);
    for (int v = 0; v < nvox; ++v) {
      const float* q = cent_normals + 3 * (size_t)v;
      const float sq = (q[0] * q[0] + q[1] * q[1]) + q[2] * q[2];
      const float len = std::sqrt(sq);
      for (int a = 0; a < 3; ++a) nn[3 * (size_t)v + a] = q[a] / len;
    }
  }
  // raw counters per histogram: 21 -> 6x6, 325 -> 325, 110 -> 7 x 5x5 + 5
  const int raw = kind == ORC_SIG_GRSD21 ? 36 : (kind == ORC_SIG_GRSD325 ? 325 : NRDIV * 25 + 5);
  std::vector<int32_t> R((size_t)hist_num * raw, 0);
  for (int v = 0; v < nvox; ++v) {
    int hist_idx = 0;
    const float* c = centroids + 3 * (size_t)v;
    if (hist_num != 1) {  // :233-246 (same text at :399-412, :565-578)
      const int tmp_x = std::floor(c[0] / leaf) - min_b[0] - off_x;
      const int tmp_y = std::floor(c[1] / leaf) - min_b[1] - off_y;
      const int tmp_z = std::floor(c[2] / leaf) - min_b[2] - off_z;
      if (tmp_x < 0 || tmp_y < 0 || tmp_z < 0) continue;
      int ix = (int)std::floor(tmp_x * inverse_subdivision_size);
      int iy = (int)std::floor(tmp_y * inverse_subdivision_size);
      int iz = (int)std::floor(tmp_z * inverse_subdivision_size);
      if (ix >= sb[0] || iy >= sb[1] || iz >= sb[2]) continue;  // see orc_grsd21: out of range in the reference
      hist_idx = ix + iy * sb[0] + iz * sb[0] * sb[1];
    }
    int32_t* H = R.data() + (size_t)hist_idx * raw;
    const int src = types[v];
    if (kind == ORC_SIG_GRSD21) {
      for (int o = 0; o < 26; ++o) {
        int nb = neighbor_centroid(c, leaf, min_b, div_b, layout, off[o]);
        int nt = (nb == -1) ? 5 : types[nb];
        H[src * 6 + nt]++;
      }
    } else if (kind == ORC_SIG_GRSD325) {
      for (int o = 0; o < 13; ++o) {  // :415-429: the 13 half offsets only
        int nb = neighbor_centroid(c, leaf, min_b, div_b, layout, off[o]);
        if (nb == -1) continue;  // "ignore EMPTY"
        H[src + types[nb] * 5 + o * 25]++;
      }
    } else {
      const float* sn = nn.data() + 3 * (size_t)v;
      if (!(std::isfinite(sn[0]) && std::isfinite(sn[1]) && std::isfinite(sn[2]))) continue;  // :583
      for (int o = 0; o < 26; ++o) {
        int nb = neighbor_centroid(c, leaf, min_b, div_b, layout, off[o]);
        if (nb == -1) {
          H[NRDIV * 25 + src]++;  // transitions_to_empty (:595-596)
          continue;
        }
        const float* m = nn.data() + 3 * (size_t)nb;
        if (std::isfinite(m[0]) && std::isfinite(m[1]) && std::isfinite(m[2])) {
          // :607-608: min(NR_DIV-1, (int) floor(sqrt(source_normal.cross(nbr).norm()) * NR_DIV));
          // cross and norm in fp32 (Eigen::Vector3f), the outer sqrt in double
          const float cx = sn[1] * m[2] - sn[2] * m[1];
          const float cy = sn[2] * m[0] - sn[0] * m[2];
          const float cz = sn[0] * m[1] - sn[1] * m[0];
          const float cn = std::sqrt((cx * cx + cy * cy) + cz * cz);
          const int bin = std::min(NRDIV - 1, (int)std::floor(std::sqrt((double)cn) * NRDIV));
          H[bin * 25 + src * 5 + types[nb]]++;
        } else {
          H[NRDIV * 25 + src]++;  // :609-610
        }
      }
    }
  }
  for (int h = 0; h < hist_num; ++h) {
    const int32_t* H = R.data() + (size_t)h * raw;
    int32_t* out = hist + (size_t)h * dim;
    int nrf = 0;
    if (kind == ORC_SIG_GRSD21) {  // :266-276
      for (int i = 0; i < 6; ++i)
        for (int j = i; j < 6; ++j) out[nrf++] = H[i * 6 + j];
    } else if (kind == ORC_SIG_GRSD325) {
      for (int i = 0; i < 325; ++i) out[i] = H[i];
    } else {  // :626-636
      for (int d = 0; d < NRDIV; ++d)
        for (int i = 0; i < NRCLASS; ++i)
          for (int j = i; j < NRCLASS; ++j) out[nrf++] = H[d * 25 + i * 5 + j];
      for (int it = 0; it < NRCLASS; ++it) out[nrf++] = H[NRDIV * 25 + it];
    }
  }
  return hist_num;
}

int orc_svm_predict(const float* features, int64_t n, int dim, int nr_class, int total_sv, double gamma,
                    const int32_t* labels, const int32_t* nr_sv, const double* rho, const double* sv_coef,
                    const double* sv, double lower, double upper, const double* fmin, const double* fmax,
                    float* out, double* dec) {
  const int npairs = nr_class * (nr_class - 1) / 2;
  std::vector<int> start(nr_class, 0);
  for (int i = 1; i < nr_class; ++i) start[i] = start[i - 1] + nr_sv[i - 1];
#pragma omp parallel for schedule(static)
  for (int64_t p = 0; p < n; ++p) {
    std::vector<double> x(dim), kvalue(total_sv), dv(npairs);
    std::vector<int> vote(nr_class, 0);
    for (int i = 0; i < dim; ++i) {
      double value = features[(size_t)p * dim + i];  // svm_classification.cpp:141
      if (fmin) {                                    // scaleFeature, svm_classification.h:68-86
        if (fmin[i] == fmax[i]) value = 0;
        else if (value <= fmin[i]) value = lower;
        else if (value >= fmax[i]) value = upper;
        else value = lower + (upper - lower) * (value - fmin[i]) / (fmax[i] - fmin[i]);
      }
      x[i] = value;
    }
    for (int s = 0; s < total_sv; ++s) {  // Kernel::k_function, RBF
      const double* y = sv + (size_t)s * dim;
      double sum = 0;
      for (int i = 0; i < dim; ++i) {
        const double d = x[i] - y[i];
        sum += d * d;
      }
      kvalue[s] = std::exp(-gamma * sum);
    }
    int pi = 0;
    for (int i = 0; i < nr_class; ++i)
      for (int j = i + 1; j < nr_class; ++j) {
        double sum = 0;
        const int si = start[i], sj = start[j], ci = nr_sv[i], cj = nr_sv[j];
        const double* coef1 = sv_coef + (size_t)(j - 1) * total_sv;
        const double* coef2 = sv_coef + (size_t)i * total_sv;
        for (int k = 0; k < ci; ++k) sum += coef1[si + k] * kvalue[si + k];
        for (int k = 0; k < cj; ++k) sum += coef2[sj + k] * kvalue[sj + k];
        sum -= rho[pi];
        dv[pi] = sum;
        if (sum > 0) ++vote[i]; else ++vote[j];
        ++pi;
      }
    int best = 0;
    for (int i = 1; i < nr_class; ++i)
      if (vote[i] > vote[best]) best = i;
    out[p] = (float)labels[best];
    if (dec)
      for (int q = 0; q < npairs; ++q) dec[(size_t)p * npairs + q] = dv[q];
  }
  return 0;
}

int orc_knn_mean_distance(const float* xyz, int n, int k, double* avg, int nthreads) {
  if (k < 2) return -1;
  KdTree tree(xyz, n);
  int nfinite = 0;
  for (int i = 0; i < n; ++i) nfinite += finite3(xyz + 3 * (size_t)i) ? 1 : 0;
  if (k > nfinite) return -2;  // noise_removal.cpp:57-62
  nthreads = resolve_threads(nthreads);
#pragma omp parallel num_threads(nthreads)
  {
    std::vector<Nb> nbs;
#pragma omp for schedule(dynamic, 256)
    for (int cp = 0; cp < n; ++cp) {
      const float* q = xyz + 3 * (size_t)cp;
      if (!finite3(q)) {
        avg[cp] = std::numeric_limits<double>::quiet_NaN();
        continue;
      }
      tree.knn(q, k, nbs);
      double a = 0.0;
      // :104-109, the first one is cp itself; sqrt on a float under `using namespace std;` (:4) is std::sqrt(float)
      for (int ni = 1; ni < k; ni++) a += (double)std::sqrt(nbs[ni].d2);
      avg[cp] = a / (k - 1);                                                // :110
    }
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// Colour half of VOSCH: rotation-invariant Color-CHLAC / C3-HLAC, 117 bins
// (color_chlac/include/color_chlac/color_chlac.hpp; called from extractC3HLACSignature117 and extractVOSCH,
// grsd_colorCHLAC_tools.hpp:787-843).  Integer colour arithmetic accumulated into float bins in the
// reference's order (voxels in cloud order, per voxel: 0th-order binary, 0th-order, then per valid
// half-stencil neighbour 1st-order binary, 1st-order) -- the float sums pass 2^24, so the order matters.
// ---------------------------------------------------------------------------------------------------

// pcl::VoxelGrid's colour of a voxel [EXTERNAL]: r, g, b unpacked, summed as floats, times 1 / count, truncated.
int orc_voxel_colors(const float* xyz, const uint32_t* rgb, int n, float leaf, uint32_t* out_rgb) {
  const float inv = 1.0f / leaf;
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  int nf = 0;
  for (int i = 0; i < n; ++i) {
    const float* p = xyz + 3 * (size_t)i;
    if (!finite3(p)) continue;
    ++nf;
    for (int a = 0; a < 3; ++a) {
      mn[a] = std::min(mn[a], p[a]);
      mx[a] = std::max(mx[a], p[a]);
    }
  }
  if (nf == 0) return 0;
  int32_t min_b[3], div_b[3];
  for (int a = 0; a < 3; ++a) {
    min_b[a] = (int32_t)std::floor(mn[a] * inv);
    div_b[a] = (int32_t)std::floor(mx[a] * inv) - min_b[a] + 1;
  }
  std::vector<std::pair<int64_t, int32_t>> keyed;
  keyed.reserve(nf);
  for (int i = 0; i < n; ++i) {
    const float* p = xyz + 3 * (size_t)i;
    if (!finite3(p)) continue;
    const int i0 = (int)(std::floor(p[0] * inv) - min_b[0]), i1 = (int)(std::floor(p[1] * inv) - min_b[1]),
              i2 = (int)(std::floor(p[2] * inv) - min_b[2]);
    keyed.emplace_back((int64_t)i0 + (int64_t)i1 * div_b[0] + (int64_t)i2 * div_b[0] * div_b[1], i);
  }
  std::sort(keyed.begin(), keyed.end());
  int nvox = 0;
  for (size_t i = 0; i < keyed.size();) {
    size_t j = i;
    float sr = 0.f, sg = 0.f, sb = 0.f;
    for (; j < keyed.size() && keyed[j].first == keyed[i].first; ++j) {
      const uint32_t c = rgb[keyed[j].second];
      sr += (float)((c >> 16) & 0xff);
      sg += (float)((c >> 8) & 0xff);
      sb += (float)(c & 0xff);
    }
    const float rcp = 1.0f / (float)(j - i);  // Eigen 3.0-3.2 `/=`: times the reciprocal (see orc_voxel_grid)
    out_rgb[nvox++] = ((uint32_t)(int)(sr * rcp) << 16) | ((uint32_t)(int)(sg * rcp) << 8) | (uint32_t)(int)(sb * rcp);
    i = j;
  }
  return nvox;
}

int orc_color_chlac117(int c3, const float* centroids, const uint32_t* rgb, int nvox, float leaf, const int32_t* min_b,
                       const int32_t* div_b, const int32_t* layout, int thR, int thG, int thB, int subdivision_size,
                       int off_x, int off_y, int off_z, int32_t* subdiv_b, float* hist) {
  if (thR < 0 || thG < 0 || thB < 0) return -2;  // color_chlac.hpp:1812-1815
  if (subdivision_size < 0) return -1;           // :203-206
  int hist_num = 1;
  float inverse_subdivision_size = 0.f;
  int sb[3] = {1, 1, 1};
  if (subdivision_size > 0) {  // setVoxelFilter, :185-201
    inverse_subdivision_size = 1.0 / subdivision_size;
    if (div_b[0] <= off_x || div_b[1] <= off_y || div_b[2] <= off_z) {
      if (subdiv_b) subdiv_b[0] = subdiv_b[1] = subdiv_b[2] = 0;
      return 0;
    }
    sb[0] = (int)std::ceil((div_b[0] - off_x) * inverse_subdivision_size);
    sb[1] = (int)std::ceil((div_b[1] - off_y) * inverse_subdivision_size);
    sb[2] = (int)std::ceil((div_b[2] - off_z) * inverse_subdivision_size);
    hist_num = sb[0] * sb[1] * sb[2];
  }
  if (subdiv_b) {
    subdiv_b[0] = sb[0];
    subdiv_b[1] = sb[1];
    subdiv_b[2] = sb[2];
  }
  if (!hist) return hist_num;
  const int DIM = 117;
  for (size_t i = 0; i < (size_t)hist_num * DIM; ++i) hist[i] = 0.f;  // :1822-1824
  int32_t off[26][3];
  fill_offsets26(off);  // the first 13 are color_chlac.h:92-113's relative_coordinates
  const float angle_norm = M_PI / 510;  // color_chlac.h:9
  // setColor (:148-166).  C3-HLAC: `255 * sin( val1 * angle_norm )` -- int * float -> float argument; the unqualified sin /
  // cos inside namespace pcl resolve to <math.h>'s double functions with the toolchain of the reference's era (GCC 4.x:
  // the float overloads live in std:: only), 255 * double -> double, truncated to int.  Only v = 255 depends on the choice
  // (254 here, 255 with sinf); the shipped feature vectors were made with the other colour coding and cannot tell.
  auto code = [&](int v, int& pos, int& neg) {
    if (c3) {
      pos = 255 * ::sin((double)(v * angle_norm));
      neg = 255 * ::cos((double)(v * angle_norm));
    } else {
      pos = v;
      neg = 255 - v;
    }
  };
  for (int v = 0; v < nvox; ++v) {
    int hist_idx = 0;
    const float* c = centroids + 3 * (size_t)v;
    if (hist_num != 1) {  // :1476-1500
      const int tmp_x = std::floor(c[0] / leaf) - min_b[0] - off_x;
      const int tmp_y = std::floor(c[1] / leaf) - min_b[1] - off_y;
      const int tmp_z = std::floor(c[2] / leaf) - min_b[2] - off_z;
      if (tmp_x < 0 || tmp_y < 0 || tmp_z < 0) continue;
      const int ix = (int)std::floor(tmp_x * inverse_subdivision_size), iy = (int)std::floor(tmp_y * inverse_subdivision_size),
                iz = (int)std::floor(tmp_z * inverse_subdivision_size);
      if (ix >= sb[0] || iy >= sb[1] || iz >= sb[2]) continue;  // see orc_grsd21: out of range in the reference
      hist_idx = ix + iy * sb[0] + iz * sb[0] * sb[1];
    }
    float* H = hist + (size_t)hist_idx * DIM;
    const uint32_t color = rgb[v];
    const int cr0 = (color >> 16) & 0xff, cg0 = (color >> 8) & 0xff, cb0 = color & 0xff;  // :1502-1505
    const int br = cr0 > thR ? 1 : 0, bg = cg0 > thG ? 1 : 0, bb = cb0 > thB ? 1 : 0;     // :129-146
    // addColorCHLAC_0_bin (:1597-1645)
    H[br ? 63 : 64]++;
    H[bg ? 65 : 66]++;
    H[bb ? 67 : 68]++;
    if (br) {
      H[bg ? 105 : 106]++;
      H[bb ? 107 : 108]++;
    } else {
      H[bg ? 109 : 110]++;
      H[bb ? 111 : 112]++;
    }
    if (bg) H[bb ? 113 : 114]++;
    else H[bb ? 115 : 116]++;
    // addColorCHLAC_0 (:1565-1595)
    int C[6];  // r, r_, g, g_, b, b_
    code(cr0, C[0], C[1]);
    code(cg0, C[2], C[3]);
    code(cb0, C[4], C[5]);
    for (int i = 0; i < 6; ++i) H[i] += C[i];
    for (int i = 0, t = 42; i < 6; ++i)
      for (int j = i; j < 6; ++j, ++t) H[t] += C[i] * C[j];
    // 13 half-stencil neighbours (:1512-1527)
    for (int o = 0; o < 13; ++o) {
      const int nb = neighbor_centroid(c, leaf, min_b, div_b, layout, off[o]);
      if (nb == -1) continue;
      const uint32_t nc = rgb[nb];
      const int r = (nc >> 16) & 0xff, g = (nc >> 8) & 0xff, b = nc & 0xff;
      // addColorCHLAC_1_bin (:1688-1743)
      const int nbr = r > thR ? 1 : 0, nbg = g > thG ? 1 : 0, nbb = b > thB ? 1 : 0;
      const int B6[6] = {nbr, 1 - nbr, nbg, 1 - nbg, nbb, 1 - nbb};
      for (int i = 0; i < 6; ++i) H[(br ? 69 : 75) + i] += B6[i];
      for (int i = 0; i < 6; ++i) H[(bg ? 81 : 87) + i] += B6[i];
      for (int i = 0; i < 6; ++i) H[(bb ? 93 : 99) + i] += B6[i];
      // addColorCHLAC_1 (:1647-1686)
      int N[6];
      code(r, N[0], N[1]);
      code(g, N[2], N[3]);
      code(b, N[4], N[5]);
      for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) H[6 + 6 * i + j] += C[i] * N[j];
    }
  }
  // normalizeColorCHLAC (:1745-1782; the constants of the two classes are equal without ENABLE_THEORY_NORMALIZATION)
  const float n_ri_0 = 1 / 255.0, n_ri_1 = 1 / 845325.0, n_1 = 1 / 65025.0, n_ri_0_bin = 1, n_ri_1_bin = 1 / 13.0, n_1_bin = 1;
  for (int h = 0; h < hist_num; ++h) {
    float* H = hist + (size_t)h * DIM;
    for (int i = 0; i < 6; ++i) H[i] *= n_ri_0;
    for (int i = 6; i < 42; ++i) H[i] *= n_ri_1;
    for (int i = 42; i < 63; ++i) H[i] *= n_1;
    for (int i = 63; i < 69; ++i) H[i] *= n_ri_0_bin;
    for (int i = 69; i < 105; ++i) H[i] *= n_ri_1_bin;
    for (int i = 105; i < DIM; ++i) H[i] *= n_1_bin;
  }
  return hist_num;
}

// cloud_geometry::nearest::extractEuclideanClusters [EXTERNAL] as called at
// cloud_tools/src/table_object_detector_passive.cpp:293,567 (nx_idx = -1: no normal test).
int orc_euclidean_clusters(const float* xyz, int n, double tolerance, int min_pts, int max_pts, int32_t* labels) {
  const float tf = (float)tolerance, r2 = tf * tf;
  CellGrid grid(xyz, n, tolerance);
  std::vector<char> processed((size_t)std::max(n, 1), 0);
  std::vector<int32_t> seed_queue;
  std::vector<Nb> nbs;
  for (int i = 0; i < n; ++i) labels[i] = -1;
  int n_clusters = 0;
  for (int i = 0; i < n; ++i) {  // seeds in index order
    if (processed[i] || !finite3(xyz + 3 * (size_t)i)) continue;
    seed_queue.clear();
    seed_queue.push_back(i);
    processed[i] = 1;
    for (size_t sq = 0; sq < seed_queue.size(); ++sq) {
      grid.query(xyz + 3 * (size_t)seed_queue[sq], r2, nbs);
      for (const Nb& nb : nbs) {
        if (processed[nb.idx]) continue;
        processed[nb.idx] = 1;
        seed_queue.push_back(nb.idx);
      }
    }
    const int size = (int)seed_queue.size();
    if (size >= min_pts && (max_pts <= 0 || size <= max_pts)) {
      for (int32_t j : seed_queue) labels[j] = n_clusters;
      ++n_clusters;
    }
  }
  return n_clusters;
}

int64_t orc_noise_filter(const double* avg, int n, double alpha, uint8_t* keep, double* mean_out, double* stddev_out) {
  // noise_removal.cpp:113-121 over the finite points
  double sum = 0, sq_sum = 0;
  int64_t m = 0;
  for (int cp = 0; cp < n; ++cp) {
    if (std::isnan(avg[cp])) continue;
    sum += avg[cp];
    sq_sum += avg[cp] * avg[cp];
    ++m;
  }
  const double mean = m ? sum / m : 0.0;
  const double variance = m ? sq_sum / m - mean * mean : 0.0;
  const double stddev = std::sqrt(variance);
  int64_t kept = 0;
  for (int cp = 0; cp < n; ++cp) {
    const bool k = !std::isnan(avg[cp]) && std::fabs(avg[cp] - mean) < alpha * stddev;  // :131
    if (keep) keep[cp] = k ? 1 : 0;
    kept += k ? 1 : 0;
  }
  if (mean_out) *mean_out = mean;
  if (stddev_out) *stddev_out = stddev;
  return kept;
}

int orc_pfh_pair(const float* ps, const float* ns, const float* pt, const float* nt, float d2, double max_dist,
                 int check_flip, int abs_angles, double* f) {
  // pfh.h:102-238, s = the query, t = the neighbour; every product below is float * double -> double
  double dP2P1[3], u[3], v[3], w[3], tmp[3];
  dP2P1[0] = pt[0] - ps[0];  // float - float (:107-109)
  dP2P1[1] = pt[1] - ps[1];
  dP2P1[2] = pt[2] - ps[2];
  double delta = std::sqrt(d2);  // pfh.cpp:246, std::sqrt(float) under `using namespace std;`
  if (delta <= 0) {              // :112-123
    double delta_sqr = dP2P1[0] * dP2P1[0] + dP2P1[1] * dP2P1[1] + dP2P1[2] * dP2P1[2];
    if (delta_sqr == 0) return 0;
    delta = std::sqrt(delta_sqr);
  }
  double angle2 = -(nt[0] * dP2P1[0] + nt[1] * dP2P1[1] + nt[2] * dP2P1[2]) / delta;  // :128-130
  const float *nsrc = ns, *ntgt = nt;
  bool do_flip = false;
  double gamma = 0;
  if (check_flip) {  // :133-140
    gamma = (ns[0] * dP2P1[0] + ns[1] * dP2P1[1] + ns[2] * dP2P1[2]) / delta;
    if (std::acos(gamma) > std::acos(angle2)) do_flip = true;
  }
  if (!check_flip || do_flip) {  // :143-152
    nsrc = nt;
    ntgt = ns;
    dP2P1[0] = -dP2P1[0];
    dP2P1[1] = -dP2P1[1];
    dP2P1[2] = -dP2P1[2];
    gamma = angle2;
  }
  if (abs_angles) gamma = std::fabs(gamma);
  u[0] = nsrc[0];
  u[1] = nsrc[1];
  u[2] = nsrc[2];
  tmp[0] = dP2P1[1] * u[2] - dP2P1[2] * u[1];  // :165-167
  tmp[1] = dP2P1[2] * u[0] - dP2P1[0] * u[2];
  tmp[2] = dP2P1[0] * u[1] - dP2P1[1] * u[0];
  double nrm = std::sqrt(tmp[0] * tmp[0] + tmp[1] * tmp[1] + tmp[2] * tmp[2]);
  if (nrm == 0) return 0;  // :171-175
  v[0] = tmp[0] / nrm;
  v[1] = tmp[1] / nrm;
  v[2] = tmp[2] / nrm;
  w[0] = u[1] * v[2] - u[2] * v[1];  // :183-185
  w[1] = u[2] * v[0] - u[0] * v[2];
  w[2] = u[0] * v[1] - u[1] * v[0];
  double beta = v[0] * ntgt[0] + v[1] * ntgt[1] + v[2] * ntgt[2];  // :191
  if (abs_angles) beta = std::fabs(beta);
  const double wy = w[0] * ntgt[0] + w[1] * ntgt[1] + w[2] * ntgt[2];
  const double ux = u[0] * ntgt[0] + u[1] * ntgt[1] + u[2] * ntgt[2];
  double alpha = abs_angles ? std::atan2(std::fabs(wy), std::fabs(ux)) : std::atan2(wy, ux);  // :200-205
  delta = delta / max_dist;  // :217
  if (abs_angles)
    alpha = alpha / (M_PI / 2);
  else {  // :221-226
    alpha = (alpha + M_PI) / (2.0 * M_PI);
    beta = (beta + 1.0) / 2.0;
    gamma = (gamma + 1.0) / 2.0;
  }
  f[0] = alpha;
  f[1] = beta;
  f[2] = gamma;
  f[3] = delta;
  return 1;
}

int orc_pfh(const float* xyz, const float* normals, int normal_stride, int n, double radius, int max_nn,
            int quantum, int flags, float* out, int nthreads) {
  if (quantum < 1) return -1;
  const bool use_dist = flags & ORC_PFH_USE_DIST, differential = flags & ORC_PFH_DIFFERENTIAL;
  const bool check_flip = flags & ORC_PFH_CHECK_FLIP, abs_angles = flags & ORC_PFH_ABS_ANGLES, average = flags & ORC_PFH_AVERAGE;
  const bool combine = flags & ORC_PFH_COMBINE;
  const int nr_features = use_dist ? 4 : 3;
  const int nr_bins = combine ? (int)std::ceil(std::pow((double)quantum, (double)nr_features)) : quantum * nr_features;  // :47-57
  // the order of the features in the n-D histogram (:113-121): fi[slot[f]] = index of feature f (alpha, beta, gamma, delta)
  const int slot4[4] = {3, 0, 2, 1}, slot3[4] = {2, 0, 1, 3};
  const int* slot = use_dist ? slot4 : slot3;
  CellGrid grid(xyz, n, radius);
  const float r2 = r2_of(radius);
  nthreads = resolve_threads(nthreads);
  std::vector<float> hist((size_t)n * nr_bins, 0.f);
  std::vector<std::vector<Nb>> lists(average ? n : 0);
#pragma omp parallel num_threads(nthreads)
  {
    std::vector<Nb> nbs;
#pragma omp for schedule(dynamic, 256)
    for (int cp = 0; cp < n; ++cp) {
      grid.query(xyz + 3 * (size_t)cp, r2, nbs);
      sort_truncate(nbs, max_nn);  // pfh.cpp:186
      float* h = hist.data() + (size_t)cp * nr_bins;
      const double npsqr = 100.0 / nbs.size();  // :212 (inf for a non-finite point, which has no neighbours at all)
      for (size_t ni = 1; ni < nbs.size(); ni++) {  // :217
        const int j = nbs[ni].idx;
        double f[4];
        if (orc_pfh_pair(xyz + 3 * (size_t)cp, normals + (size_t)normal_stride * cp, xyz + 3 * (size_t)j,
                         normals + (size_t)normal_stride * j, nbs[ni].d2, 2 * radius, check_flip, abs_angles, f)) {
          if (combine) {  // :239-258: the feature indices as the digits of a number in base quantum
            int fi[4] = {0, 0, 0, 0};
            for (int ft = 0; ft < nr_features; ++ft)
              fi[slot[ft]] = std::max(0, std::min(quantum - 1, (int)std::floor(quantum * f[ft])));
            int index = 0, power = 1;
            for (int d = 0; d < nr_features; ++d) {
              index += power * fi[d];
              power *= quantum;
            }
            h[index] += npsqr;
          } else {
            for (int ft = 0; ft < nr_features; ++ft) {  // :224-228, :267-271 with a_, b_, c_, d_ = 0, 1, 2, 3
              const int fi = std::max(0, std::min(quantum - 1, (int)std::floor(quantum * f[ft])));
              h[ft * quantum + fi] += npsqr;  // float += double
            }
          }
        } else if (combine) {
          for (int i = 0; i < nr_bins; i++) h[i] += npsqr / nr_bins;  // :279-281
        } else {
          for (int i = 0; i < nr_bins; i++) h[i] += npsqr / quantum;  // :284-286
        }
      }
      if (average) lists[cp] = nbs;
    }
  }
#pragma omp parallel for schedule(dynamic, 256) num_threads(nthreads)
  for (int cp = 0; cp < n; ++cp) {
    float* o = out + (size_t)cp * nr_bins;
    if (!average) {
      for (int b = 0; b < nr_bins; ++b) o[b] = hist[(size_t)cp * nr_bins + b];  // :295-300
    } else {
      for (int b = 0; b < nr_bins; ++b) o[b] = 0.f;
      const std::vector<Nb>& nbs = lists[cp];
      if (!nbs.empty()) {  // :311-312
        double sum_weight = 0.0;
        for (size_t ni = 1; ni < nbs.size(); ni++) {  // :317-327
          const double weight = 1.0 / nbs[ni].d2;
          sum_weight += weight;
          const float* hn = hist.data() + (size_t)nbs[ni].idx * nr_bins;
          for (int b = 0; b < nr_bins; ++b) o[b] += hn[b] * weight;
        }
        for (int b = 0; b < nr_bins; ++b) o[b] /= sum_weight;  // :330-331
      }
    }
    if (differential && !combine)  // :337-350 (:345 !combine_ && differential_)
      for (int ft = 0; ft < nr_features; ++ft)
        for (int b = quantum - 1; b > 0; b--) o[ft * quantum + b] -= o[ft * quantum + b - 1];
  }
  return nr_bins;
}


// fitSACPlane, cloud_tools/src/table_object_detector_passive.cpp:621-659: sample_consensus::MSAC over SACModelPlane
// [point_cloud_mapping, EXTERNAL: not in /root/reference -- parity unpinned; the published algorithm restated:
// MSAC::computeModel's loop (penalty = sum min(distance, threshold); k = log(1 - p) / log(1 - w^3) after every
// improvement; iterations_ > max_iterations_ ends it), SACModelPlane::computeModelCoefficients (plane through three
// points), getDistancesToModel (|a x + b y + c z + d| in double), refineCoefficients (computePointNormal: centroid,
// covariance, eigenvector of the smallest eigenvalue), selectWithinDistance, projectPointsInPlace].  The sample sequence
// is given (triples: positions into the index list); a degenerate triple is skipped, as getSamples draws again.
int64_t orc_fit_plane_msac(const float* xyz, int64_t n, const int32_t* indices, int64_t n_idx, double threshold,
                           int32_t max_iterations, double probability, const int32_t* triples, int64_t n_triples,
                           double coeff[4], int32_t* inliers, float* projected_xyz, int32_t* iterations_run,
                           int32_t* best_iteration) {
  const int64_t m = indices ? n_idx : n;
  auto point = [&](int64_t pos) { return xyz + 3 * (size_t)(indices ? indices[pos] : pos); };
  for (int k = 0; k < 4; ++k) coeff[k] = 0.0;
  *iterations_run = 0;
  *best_iteration = -1;
  if (m < 3) return 0;
  double best_penalty = DBL_MAX, k = 1.0, best[4] = {0, 0, 0, 0};
  int iterations = 0;
  for (int64_t smp = 0; iterations < k && smp < n_triples; ++smp) {
    const int32_t* t = triples + 3 * (size_t)smp;
    double mdl[4];
    bool ok = t[0] != t[1] && t[0] != t[2] && t[1] != t[2];
    if (ok) {
      const float *p0 = point(t[0]), *p1 = point(t[1]), *p2 = point(t[2]);
      const double ux = (double)p1[0] - p0[0], uy = (double)p1[1] - p0[1], uz = (double)p1[2] - p0[2];
      const double vx = (double)p2[0] - p0[0], vy = (double)p2[1] - p0[1], vz = (double)p2[2] - p0[2];
      double a = uy * vz - uz * vy, b = uz * vx - ux * vz, c = ux * vy - uy * vx;
      const double len = std::sqrt(a * a + b * b + c * c);
      ok = len > 0.0 && std::isfinite(len);
      if (ok) {
        a /= len;
        b /= len;
        c /= len;
        mdl[0] = a;
        mdl[1] = b;
        mdl[2] = c;
        mdl[3] = -(a * (double)p0[0] + b * (double)p0[1] + c * (double)p0[2]);
      }
    }
    if (!ok) continue;  // getSamples draws again: a degenerate triple is not an iteration
    {
      double penalty = 0;
      int64_t count = 0;
      for (int64_t i = 0; i < m; ++i) {
        const float* p = point(i);
        const double d = std::fabs(mdl[0] * (double)p[0] + mdl[1] * (double)p[1] + mdl[2] * (double)p[2] + mdl[3]);
        penalty += std::min(d, threshold);
        count += d <= threshold ? 1 : 0;
      }
      if (penalty < best_penalty) {
        best_penalty = penalty;
        std::memcpy(best, mdl, sizeof(best));
        *best_iteration = (int32_t)smp;
        const double w = (double)count / (double)m;
        double p_no_outliers = 1.0 - std::pow(w, 3.0);
        p_no_outliers = std::max(std::numeric_limits<double>::epsilon(), p_no_outliers);
        p_no_outliers = std::min(1.0 - std::numeric_limits<double>::epsilon(), p_no_outliers);
        k = std::log(1.0 - probability) / std::log(p_no_outliers);
      }
    }
    iterations += 1;
    if (iterations > max_iterations) break;
  }
  *iterations_run = iterations;
  if (*best_iteration < 0) return 0;
  // refineCoefficients over the best model's inliers
  double sx = 0, sy = 0, sz = 0, cnt = 0;
  std::vector<char> in((size_t)m, 0);
  for (int64_t i = 0; i < m; ++i) {
    const float* p = point(i);
    const double d = std::fabs(best[0] * (double)p[0] + best[1] * (double)p[1] + best[2] * (double)p[2] + best[3]);
    if (d <= threshold) {
      in[(size_t)i] = 1;
      sx += p[0];
      sy += p[1];
      sz += p[2];
      cnt += 1;
    }
  }
  double ref[4] = {best[0], best[1], best[2], best[3]};
  if (cnt >= 3) {
    const double cx = sx / cnt, cy = sy / cnt, cz = sz / cnt;
    double cov[6] = {0, 0, 0, 0, 0, 0};
    for (int64_t i = 0; i < m; ++i) {
      if (!in[(size_t)i]) continue;
      const float* p = point(i);
      const double dx = (double)p[0] - cx, dy = (double)p[1] - cy, dz = (double)p[2] - cz;
      cov[0] += dx * dx;
      cov[1] += dx * dy;
      cov[2] += dx * dz;
      cov[3] += dy * dy;
      cov[4] += dy * dz;
      cov[5] += dz * dz;
    }
    double w[3], v[3][3];
    eig3_jacobi(cov, w, v);
    double nx = v[0][0], ny = v[1][0], nz = v[2][0];
    const double len = std::sqrt(nx * nx + ny * ny + nz * nz);
    nx /= len;
    ny /= len;
    nz /= len;
    if (std::isfinite(nx) && std::isfinite(ny) && std::isfinite(nz)) {
      if (nx * best[0] + ny * best[1] + nz * best[2] < 0) {  // the sign is arbitrary upstream: the sampled model's side
        nx = -nx;
        ny = -ny;
        nz = -nz;
      }
      ref[0] = nx;
      ref[1] = ny;
      ref[2] = nz;
      ref[3] = -(nx * cx + ny * cy + nz * cz);
    }
  }
  std::memcpy(coeff, ref, sizeof(ref));
  int64_t n_in = 0;
  for (int64_t i = 0; i < m; ++i) {
    const float* p = point(i);
    const double dist = ref[0] * (double)p[0] + ref[1] * (double)p[1] + ref[2] * (double)p[2] + ref[3];
    if (std::fabs(dist) <= threshold) {
      if (inliers) inliers[n_in] = (int32_t)(indices ? indices[i] : i);
      if (projected_xyz) {
        projected_xyz[3 * n_in] = (float)((double)p[0] - dist * ref[0]);
        projected_xyz[3 * n_in + 1] = (float)((double)p[1] - dist * ref[1]);
        projected_xyz[3 * n_in + 2] = (float)((double)p[2] - dist * ref[2]);
      }
      ++n_in;
    }
  }
  return n_in;
}

}  // extern "C"
