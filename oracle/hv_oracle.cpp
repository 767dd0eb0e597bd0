/*
 * hv_oracle.cpp — CPU restatement of pcl::GlobalHypothesesVerification as the reference calls it
 * (/root/reference/SHOT_hypothesis.cpp:631-653: setSceneCloud, addModels(instances, true), the seven
 * set* calls, verify, getMask).
 *
 * TEST INFRASTRUCTURE ONLY (see pcl_oracle.h).  PARITY UNPINNED: PCL is neither vendored in the
 * reference nor installed here; this follows PCL 1.8's recognition/hv/hypotheses_verification.h,
 * recognition/impl/hv/hv_go.hpp, recognition/impl/hv/occlusion_reasoning.hpp and the bundled metslib
 * (simulated-annealing.hh, termination-criteria.hh) as recalled.  Choices that cannot be checked here:
 *   - the search tree over the down-sampled scene is taken to index the cloud AFTER the NaN-normal
 *     compaction (PCL builds it before and rebuilds it only when detect_clutter_ is set; used as is, its
 *     indices would run past explained_by_RM_, which is undefined behaviour — nothing to restate);
 *   - the uniform variate of the annealing acceptance test: boost::uniform_real over boost::mt19937
 *     (x / 2^32, mode 0) or std::tr1::uniform_real called on the raw engine (the 32-bit integer itself,
 *     so only improving moves are ever accepted, mode 1);
 *   - the self-occlusion depth map is 75 x 75 with a 5 mm margin (literals in addModels);
 *   - detect_clutter_ = true (smooth-surface segmentation + clutter cue) is not restated; the reference
 *     sets it false (SHOT_hypothesis.cpp:64, :647).
 * Building blocks come from pcl_oracle.cpp through its C API (VoxelGrid, NormalEstimation, radiusSearch).
 */
#include "pcl_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <map>
#include <random>
#include <vector>

namespace {

const float kNaN = std::numeric_limits<float>::quiet_NaN();

/* pcl::occlusion_reasoning::ZBuffering (occlusion_reasoning.hpp) */
struct ZBuffer {
  int cx_, cy_;
  float f_;
  std::vector<float> depth_;
  ZBuffer(int rx, int ry, float f) : cx_(rx), cy_(ry), f_(f) {}

  /* int u = static_cast<int>(f_ * x / z + cx): a non-finite or out-of-range value converts to INT_MIN on
   * x86 (cvttss2si) and fails the u < 0 test */
  bool pixel(float x, float y, float z, int &u, int &v) const {
    const float cx = static_cast<float>(cx_) / 2.f - 0.5f;
    const float cy = static_cast<float>(cy_) / 2.f - 0.5f;
    const float fu = f_ * x / z + cx;
    const float fv = f_ * y / z + cy;
    if (!(fu > -2147483648.f && fu < 2147483648.f) || !(fv > -2147483648.f && fv < 2147483648.f)) return false;
    u = static_cast<int>(fu);
    v = static_cast<int>(fv);
    return !(u >= cx_ || v >= cy_ || u < 0 || v < 0);
  }

  /* computeDepthMap(cloud, compute_focal = true) */
  void compute_depth_map(const float *pts, int n, int stride) {
    const float cx = static_cast<float>(cx_) / 2.f - 0.5f;
    float max_u, max_v, min_u, min_v;
    max_u = max_v = std::numeric_limits<float>::max() * -1;
    min_u = min_v = std::numeric_limits<float>::max();
    for (int i = 0; i < n; ++i) {
      const float *p = pts + (size_t)i * stride;
      const float b_x = p[0] / p[2];
      if (b_x > max_u) max_u = b_x;
      if (b_x < min_u) min_u = b_x;
      const float b_y = p[1] / p[2];
      if (b_y > max_v) max_v = b_y;
      if (b_y < min_v) min_v = b_y;
    }
    const float maxC = std::max(std::max(std::abs(max_u), std::abs(max_v)), std::max(std::abs(min_u), std::abs(min_v)));
    f_ = cx / maxC;
    depth_.assign((size_t)cx_ * cy_, kNaN);
    for (int i = 0; i < n; ++i) {
      const float *p = pts + (size_t)i * stride;
      int u, v;
      if (!pixel(p[0], p[1], p[2], u, v)) continue;
      float &d = depth_[u + v * cx_];
      if ((p[2] < d) || (!std::isfinite(d))) d = p[2];
    }
  }

  /* filter(model, indices_to_keep, thres) */
  bool keeps(const float *p, float thres) const {
    int u, v;
    if (!pixel(p[0], p[1], p[2], u, v)) return false;
    const float d = depth_[u + v * cx_];
    if ((p[2] - thres) > d || !std::isfinite(d)) return false;
    return true;
  }
};

/* glibc rand(): what std::random_shuffle draws from in libstdc++ */
struct GlibcRand {
  uint32_t st[31];
  explicit GlibcRand(unsigned seed) {
    if (seed == 0) seed = 1;
    std::vector<uint32_t> r(344);
    int32_t word = (int32_t)seed;
    r[0] = (uint32_t)word;
    for (int i = 1; i < 31; ++i) {
      const int64_t hi = word / 127773, lo = word % 127773;
      int64_t w = 16807 * lo - 2836 * hi;
      if (w < 0) w += 2147483647;
      word = (int32_t)w;
      r[i] = (uint32_t)word;
    }
    for (int i = 31; i < 34; ++i) r[i] = r[i - 31];
    for (int i = 34; i < 344; ++i) r[i] = r[i - 31] + r[i - 3];
    for (int i = 0; i < 31; ++i) st[i] = r[313 + i];
  }
  int next() {
    const uint32_t o = st[0] + st[28];
    for (int i = 0; i < 30; ++i) st[i] = st[i + 1];
    st[30] = o;
    return (int)(o >> 1);
  }
};

struct Lists {
  std::vector<int> expl_off, expl_idx, occ_off, occ_idx;
  std::vector<float> expl_w;
};
thread_local Lists g_last;

/* GlobalHypothesesVerification::SAOptimize / evaluateSolution / updateExplainedVector / updateCMDuplicity and
 * mets::simulated_annealing::search with move_manager (one flip move per hypothesis, std::random_shuffle per
 * iteration), best_ever_solution, noimprove_termination_criteria(max_iterations), linear_cooling (0.95),
 * starting temperature initial_temp, stop temperature 1e-7, K = 2 */
struct Annealer {
  int H, ns;
  const int *eo, *ei, *oo, *oi;
  const float *ew;
  const float *outliers_weight;
  const int *bad_information;
  float w_cm;
  std::vector<int> explained, occupancy;
  std::vector<float> weighted;
  float previous_explained_value = 0.f, previous_bad_info = 0.f, previous_unexplained = 0.f;
  int previous_duplicity = 0, previous_duplicity_cm = 0;

  void update_explained(int h, float sign) {
    float add_to_explained = 0.f;
    int add_to_duplicity = 0;
    for (int k = eo[h]; k < eo[h + 1]; ++k) {
      const int j = ei[k];
      const bool prev_dup = explained[j] > 1;
      explained[j] += static_cast<int>(sign);
      weighted[j] += ew[k] * sign;
      add_to_explained += ew[k];
      if ((explained[j] > 1) && prev_dup)
        add_to_duplicity += static_cast<int>(sign);
      else if ((explained[j] == 1) && prev_dup)
        add_to_duplicity -= 2;
      else if ((explained[j] > 1) && !prev_dup)
        add_to_duplicity += 2;
    }
    previous_explained_value += add_to_explained * sign;
    previous_duplicity += add_to_duplicity;
  }
  void update_cm(int h, float sign) {
    int add = 0;
    for (int k = oo[h]; k < oo[h + 1]; ++k) {
      const int c = oi[k];
      const bool prev_dup = occupancy[c] > 1;
      occupancy[c] += static_cast<int>(sign);
      if ((occupancy[c] > 1) && prev_dup)
        add += static_cast<int>(sign);
      else if ((occupancy[c] == 1) && prev_dup)
        add -= 2;
      else if ((occupancy[c] > 1) && !prev_dup)
        add += 2;
    }
    previous_duplicity_cm += add;
  }
  double evaluate(const std::vector<char> &active, int changed) {
    float sign = 1.f;
    if (active[changed]) {
      update_explained(changed, 1.f);
      update_cm(changed, 1.f);
    } else {
      update_explained(changed, -1.f);
      update_cm(changed, -1.f);
      sign = -1.f;
    }
    const int duplicity = previous_duplicity;
    const float good_info = previous_explained_value;
    const float unexplained_info = previous_unexplained;
    const float bad_info =
        static_cast<float>(previous_bad_info) + (outliers_weight[changed] * static_cast<float>(bad_information[changed])) * sign;
    previous_bad_info = bad_info;
    int n_active_hyp = 0;
    for (size_t i = 0; i < active.size(); ++i)
      if (active[i]) n_active_hyp++;
    const float duplicity_cm = static_cast<float>(previous_duplicity_cm) * w_cm;
    return static_cast<double>((good_info - bad_info - static_cast<float>(duplicity) - unexplained_info - duplicity_cm -
                                static_cast<float>(n_active_hyp)) *
                               -1.f);
  }
};

int anneal(int H, int ns, const int *eo, const int *ei, const float *ew, const int *oo, const int *oi, int n_cells,
           const float *outliers_weight, const int *bad_information, const orc_hv_params *P, unsigned char *mask,
           double *best_cost_out, int *accepted_out) {
  Annealer A;
  A.H = H, A.ns = ns, A.eo = eo, A.ei = ei, A.ew = ew, A.oo = oo, A.oi = oi;
  A.outliers_weight = outliers_weight, A.bad_information = bad_information, A.w_cm = P->w_occupied_multiple_cm;
  A.explained.assign((size_t)ns, 0);
  A.weighted.assign((size_t)ns, 0.f);
  A.occupancy.assign((size_t)n_cells, 0);
  std::vector<char> solution((size_t)H, 1); /* verify(): subsolution(cc.size(), true) */
  for (int h = 0; h < H; ++h)
    for (int k = oo[h]; k < oo[h + 1]; ++k) A.occupancy[oi[k]]++;
  int occupied_multiple = 0;
  for (int c = 0; c < n_cells; ++c)
    if (A.occupancy[c] > 1) occupied_multiple += A.occupancy[c];
  A.previous_duplicity_cm = occupied_multiple;
  for (int h = 0; h < H; ++h)
    for (int k = eo[h]; k < eo[h + 1]; ++k) {
      A.explained[ei[k]]++;
      A.weighted[ei[k]] += ew[k];
    }
  /* getTotalExplainedInformation */
  float good_information = 0;
  int duplicity = 0;
  for (int i = 0; i < ns; ++i) {
    if (A.explained[i] > 0) good_information += A.weighted[i];
    if (A.explained[i] > 1) duplicity += A.explained[i];
  }
  float bad_information_sum = 0;
  const float unexplained = 0.f; /* getUnexplainedInformationInNeighborhood: every weight is 0 without the clutter cue */
  for (int h = 0; h < H; ++h)
    if (solution[h]) bad_information_sum += outliers_weight[h] * static_cast<float>(bad_information[h]);
  A.previous_explained_value = good_information;
  A.previous_duplicity = duplicity;
  A.previous_bad_info = bad_information_sum;
  A.previous_unexplained = unexplained;
  double cost = static_cast<double>((good_information - bad_information_sum - static_cast<float>(duplicity) -
                                     static_cast<float>(occupied_multiple) * P->w_occupied_multiple_cm -
                                     static_cast<float>(H) - unexplained) *
                                    -1.f);
  std::vector<char> best = solution;
  double best_cost = cost;
  int accepted = 0;

  std::vector<int> moves((size_t)H);
  for (int i = 0; i < H; ++i) moves[i] = i;
  GlibcRand rnd(P->rand_seed);
  std::mt19937 rng(P->mt_seed);
  /* noimprove_termination_criteria */
  double crit_best = std::numeric_limits<double>::max();
  int iterations_left = P->max_iterations;
  double temp = P->initial_temp;
  const double K = 2.0;
  for (;;) {
    /* termination_criteria_m(working_solution) && temperature test */
    if (cost < crit_best - 1e-7) {
      crit_best = cost;
      iterations_left = P->max_iterations;
    }
    if (iterations_left <= 0) break;
    --iterations_left;
    if (!(temp > 1e-7)) break;
    const double actual_cost = cost;
    /* move_manager::refresh: std::random_shuffle */
    for (int i = 1; i < H; ++i) {
      const int j = rnd.next() % (i + 1);
      if (i != j) std::swap(moves[i], moves[j]);
    }
    for (int mi = 0; mi < H; ++mi) {
      const int m = moves[mi];
      solution[m] = !solution[m];
      cost = A.evaluate(solution, m);
      const double delta = cost - actual_cost;
      bool take = delta < 0;
      if (!take) {
        const uint32_t x = (uint32_t)rng();
        const double u = P->sa_uniform_mode == 1 ? (double)x : (double)x / 4294967296.0;
        take = u < std::exp(K * -delta / temp);
      }
      if (take) {
        if (cost < best_cost) {
          best_cost = cost;
          best = solution;
        }
        accepted++;
        break;
      }
      solution[m] = !solution[m];
      cost = A.evaluate(solution, m);
    }
    temp *= 0.95;
  }
  for (int h = 0; h < H; ++h) mask[h] = best[h] ? 1 : 0;
  if (best_cost_out) *best_cost_out = best_cost;
  if (accepted_out) *accepted_out = accepted;
  return 0;
}

}  // namespace

extern "C" {

void orc_hv_params_default(orc_hv_params *p) {
  /* HypothesisVerification / GlobalHypothesesVerification constructors */
  p->resolution = 0.005f;
  p->inlier_threshold = 0.005f; /* inliers_threshold_ = resolution_ */
  p->occlusion_threshold = 0.005f;
  p->regularizer = 1.f;
  p->radius_normals = 0.01f;
  p->res_occupancy_grid = 0.01f;
  p->w_occupied_multiple_cm = 4.f;
  p->initial_temp = 1000.f;
  p->max_iterations = 5000;
  p->occlusion_reasoning = 0;
  p->zbuffer_scene_resolution = 100;
  p->zbuffer_self_resolution = 75;
  p->self_occlusion_threshold = 0.005f;
  p->detect_clutter = 1;
  p->radius_clutter = 0.03f;
  p->clutter_regularizer = 5.f;
  p->rand_seed = 1u;
  p->mt_seed = 5489u;
  p->sa_uniform_mode = 0;
}

int orc_hv_optimize(int H, int ns, const int *expl_off, const int *expl_idx, const float *expl_w, const int *occ_off,
                    const int *occ_idx, int n_cells, const float *outliers_weight, const int *bad_information,
                    const orc_hv_params *P, unsigned char *mask, double *best_cost, int *accepted_moves) {
  if (H <= 0) return 0;
  return anneal(H, ns, expl_off, expl_idx, expl_w, occ_off, occ_idx, n_cells, outliers_weight, bad_information, P, mask,
                best_cost, accepted_moves);
}

int orc_hv_verify(const float *scene, int n, int sstride, const float *models, const int *model_offsets, int H, int mstride,
                  const orc_hv_params *P, unsigned char *mask, orc_hv_info *info, double *best_cost, int *accepted_moves,
                  int *n_scene_points, int *n_cells_out) {
  if (P->detect_clutter) return -1;
  for (int h = 0; h < H; ++h) mask[h] = 0;
  std::vector<orc_hv_info> inf((size_t)H);
  memset(inf.data(), 0, sizeof(orc_hv_info) * (size_t)H);
  /* ---- setSceneCloud: VoxelGrid(resolution_) */
  std::vector<float> S0((size_t)std::max(n, 1) * 3);
  int n0 = orc_voxel_grid(scene, n, sstride, P->resolution, P->resolution, P->resolution, S0.data());
  if (n0 < 0) return -2;
  /* ---- addModels(models, occlusion_reasoning) */
  std::vector<std::vector<float>> visible((size_t)H);
  ZBuffer zscene(P->zbuffer_scene_resolution, P->zbuffer_scene_resolution, 1.f);
  if (P->occlusion_reasoning) zscene.compute_depth_map(scene, n, sstride);
  for (int h = 0; h < H; ++h) {
    const float *mp = models + (size_t)model_offsets[h] * mstride;
    const int nm = model_offsets[h + 1] - model_offsets[h];
    std::vector<float> &vis = visible[h];
    if (!P->occlusion_reasoning) {
      for (int i = 0; i < nm; ++i) vis.insert(vis.end(), mp + (size_t)i * mstride, mp + (size_t)i * mstride + 3);
    } else {
      ZBuffer zself(P->zbuffer_self_resolution, P->zbuffer_self_resolution, 1.f);
      zself.compute_depth_map(mp, nm, mstride);
      for (int i = 0; i < nm; ++i) {
        const float *p = mp + (size_t)i * mstride;
        if (!zself.keeps(p, P->self_occlusion_threshold)) continue;
        if (!zscene.keeps(p, P->occlusion_threshold)) continue;
        vis.insert(vis.end(), p, p + 3);
      }
    }
    inf[h].n_visible = (int)(vis.size() / 3);
  }
  /* ---- initialize(): scene normals, NaN compaction */
  std::vector<float> N0((size_t)std::max(n0, 1) * 4);
  if (n0 > 0) orc_normals(S0.data(), n0, 3, S0.data(), n0, 3, 0, P->radius_normals, nullptr, N0.data());
  std::vector<float> S, SN;
  for (int i = 0; i < n0; ++i) {
    const float *nn = &N0[(size_t)i * 4];
    if (!std::isfinite(nn[0]) || !std::isfinite(nn[1]) || !std::isfinite(nn[2])) continue;
    S.insert(S.end(), &S0[(size_t)i * 3], &S0[(size_t)i * 3] + 3);
    SN.insert(SN.end(), nn, nn + 3);
  }
  const int ns = (int)(S.size() / 3);
  if (n_scene_points) *n_scene_points = ns;
  /* ---- addModel per hypothesis */
  Lists &L = g_last;
  L = Lists();
  L.expl_off.push_back(0);
  L.occ_off.push_back(0);
  std::vector<int> indices;
  std::vector<float> outliers_weight;
  std::vector<int> bad_information;
  for (int h = 0; h < H; ++h) {
    const std::vector<float> &vis = visible[h];
    const int nv = (int)(vis.size() / 3);
    std::vector<float> V((size_t)std::max(nv, 1) * 3);
    int nvox = nv > 0 ? orc_voxel_grid(vis.data(), nv, 3, P->resolution, P->resolution, P->resolution, V.data()) : 0;
    if (nvox < 0) return -2;
    /* (VoxelGrid output has no NaN rows) */
    if (nvox <= 0) continue; /* "The model cloud has no points.." */
    std::vector<float> VN((size_t)nvox * 4);
    orc_normals(V.data(), nvox, 3, V.data(), nvox, 3, 0, P->radius_normals, nullptr, VN.data());
    std::vector<float> M, MN;
    for (int i = 0; i < nvox; ++i) {
      const float *nn = &VN[(size_t)i * 4];
      if (!std::isfinite(nn[0]) || !std::isfinite(nn[1]) || !std::isfinite(nn[2])) continue;
      M.insert(M.end(), &V[(size_t)i * 3], &V[(size_t)i * 3] + 3);
      MN.insert(MN.end(), nn, nn + 3);
    }
    const int nm = (int)(M.size() / 3);
    inf[h].valid = 1;
    inf[h].n_points = nm;
    std::vector<int64_t> off((size_t)nm + 1, 0);
    std::vector<int> nidx;
    std::vector<float> nd2;
    if (nm > 0 && ns > 0) {
      const int64_t total =
          orc_radius_search(S.data(), ns, 3, M.data(), nm, 3, P->inlier_threshold, off.data(), nullptr, nullptr, 0);
      nidx.resize((size_t)std::max<int64_t>(total, 1));
      nd2.resize((size_t)std::max<int64_t>(total, 1));
      orc_radius_search(S.data(), ns, 3, M.data(), nm, 3, P->inlier_threshold, off.data(), nidx.data(), nd2.data(), total);
    }
    std::map<int, std::vector<std::pair<int, float>>> model_explains_scene_points;
    int o = 0;
    for (int i = 0; i < nm; ++i) {
      if (off[i + 1] == off[i]) {
        o++;
      } else {
        for (int64_t k = off[i]; k < off[i + 1]; ++k) model_explains_scene_points[nidx[k]].push_back(std::make_pair(i, nd2[k]));
      }
    }
    /* outliers_weight_ = accumulate(o copies of regularizer_) / o; 1 when there is no outlier */
    float acc = 0.f;
    for (int i = 0; i < o; ++i) acc += P->regularizer;
    float ow = acc / static_cast<float>(o);
    if (o == 0) ow = 1.f;
    float sum = 0.f;
    for (auto &kv : model_explains_scene_points) {
      size_t closest = 0;
      float min_d = std::numeric_limits<float>::min();
      for (size_t i = 0; i < kv.second.size(); ++i)
        if (kv.second[i].second > min_d) {
          min_d = kv.second[i].second;
          closest = i;
        }
      if (orc_get_variant() & 2048u) { /* sensitivity variant: the model point that IS closest (what the comment in PCL says) */
        closest = 0;
        for (size_t i = 1; i < kv.second.size(); ++i)
          if (kv.second[i].second < kv.second[closest].second) closest = i;
      }
      const float d = kv.second[closest].second;
      const float d_weight = -(d * d / (P->inlier_threshold)) + 1;
      const float *sn = &SN[(size_t)kv.first * 3];
      const float *mn = &MN[(size_t)kv.second[closest].first * 3];
      float dotp = ((sn[0] * mn[0] + sn[1] * mn[1]) + sn[2] * mn[2]) * 1.f;
      if (dotp < 0.f) dotp = 0.f;
      L.expl_idx.push_back(kv.first);
      L.expl_w.push_back(d_weight * dotp);
      sum += d_weight * dotp;
    }
    L.expl_off.push_back((int)L.expl_idx.size());
    inf[h].n_outliers = o;
    inf[h].outliers_weight = ow;
    inf[h].n_explained = (int)model_explains_scene_points.size();
    inf[h].explained_sum = sum;
    indices.push_back(h);
    outliers_weight.push_back(ow);
    bad_information.push_back(o);
  }
  const int Hv = (int)indices.size();
  /* ---- occupancy grid of the complete models */
  int n_cells = 0;
  if (Hv > 0) {
    float mn[3], mx[3];
    mn[0] = mn[1] = mn[2] = std::numeric_limits<float>::max();
    mx[0] = mx[1] = mx[2] = (std::numeric_limits<float>::max() - 0.001f) * -1;
    for (int v = 0; v < Hv; ++v) {
      const int h = indices[v];
      for (int i = model_offsets[h]; i < model_offsets[h + 1]; ++i) {
        const float *p = models + (size_t)i * mstride;
        if (!std::isfinite(p[0]) || !std::isfinite(p[1]) || !std::isfinite(p[2])) continue;
        for (int a = 0; a < 3; ++a) {
          if (p[a] < mn[a]) mn[a] = p[a];
          if (p[a] > mx[a]) mx[a] = p[a];
        }
      }
    }
    const float res = P->res_occupancy_grid;
    const int size_x = static_cast<int>(std::ceil(std::abs(mx[0] - mn[0]) / res)) + 1;
    const int size_y = static_cast<int>(std::ceil(std::abs(mx[1] - mn[1]) / res)) + 1;
    const int size_z = static_cast<int>(std::ceil(std::abs(mx[2] - mn[2]) / res)) + 1;
    if ((int64_t)size_x * size_y * size_z > (int64_t)1 << 27) return -3;
    n_cells = size_x * size_y * size_z;
    for (int v = 0; v < Hv; ++v) {
      const int h = indices[v];
      std::map<int, bool> banned;
      for (int i = model_offsets[h]; i < model_offsets[h + 1]; ++i) {
        const float *p = models + (size_t)i * mstride;
        if (!std::isfinite(p[0]) || !std::isfinite(p[1]) || !std::isfinite(p[2])) continue;
        const int pos_x = static_cast<int>(std::floor((p[0] - mn[0]) / res));
        const int pos_y = static_cast<int>(std::floor((p[1] - mn[1]) / res));
        const int pos_z = static_cast<int>(std::floor((p[2] - mn[2]) / res));
        const int idx = pos_z * size_x * size_y + pos_y * size_x + pos_x;
        if (banned.find(idx) == banned.end()) {
          L.occ_idx.push_back(idx);
          banned[idx] = true;
        }
      }
      L.occ_off.push_back((int)L.occ_idx.size());
      inf[h].n_occupancy = L.occ_off[v + 1] - L.occ_off[v];
    }
  }
  if (n_cells_out) *n_cells_out = n_cells;
  if (info) memcpy(info, inf.data(), sizeof(orc_hv_info) * (size_t)H);
  if (best_cost) *best_cost = 0.0;
  if (accepted_moves) *accepted_moves = 0;
  if (Hv == 0) return 0;
  std::vector<unsigned char> sub((size_t)Hv);
  anneal(Hv, ns, L.expl_off.data(), L.expl_idx.data(), L.expl_w.data(), L.occ_off.data(), L.occ_idx.data(), n_cells,
         outliers_weight.data(), bad_information.data(), P, sub.data(), best_cost, accepted_moves);
  for (int v = 0; v < Hv; ++v) mask[indices[v]] = sub[v];
  return 0;
}

/* lists of the last orc_hv_verify on this thread: which = 0 expl_off, 1 expl_idx, 2 expl_w (float), 3 occ_off, 4 occ_idx */
int orc_hv_last_size(int which) {
  const Lists &L = g_last;
  switch (which) {
    case 0: return (int)L.expl_off.size();
    case 1: return (int)L.expl_idx.size();
    case 2: return (int)L.expl_w.size();
    case 3: return (int)L.occ_off.size();
    case 4: return (int)L.occ_idx.size();
  }
  return -1;
}
int orc_hv_last_copy(int which, void *dst) {
  const Lists &L = g_last;
  switch (which) {
    case 0: memcpy(dst, L.expl_off.data(), L.expl_off.size() * 4); return 0;
    case 1: memcpy(dst, L.expl_idx.data(), L.expl_idx.size() * 4); return 0;
    case 2: memcpy(dst, L.expl_w.data(), L.expl_w.size() * 4); return 0;
    case 3: memcpy(dst, L.occ_off.data(), L.occ_off.size() * 4); return 0;
    case 4: memcpy(dst, L.occ_idx.data(), L.occ_idx.size() * 4); return 0;
  }
  return -1;
}

} /* extern "C" */
