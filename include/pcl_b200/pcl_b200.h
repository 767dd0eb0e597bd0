/*
 * pcl_b200.h — PCL-style, header-only C++ adapters over the C ABI of libb200reg.so (b200reg.h).
 *
 * The reference programs reach the recognition hot path only through a handful of PCL classes
 * (SURVEY.md §8(b)).  This header provides classes with the same names, setters and compute /
 * search / recognize signatures, so the body of a reference cloud_cb() (e.g. SHOT.cpp:298-482,
 * FPFH_demo.cpp:405-538) compiles against it with
 *
 *     #include <pcl_b200/pcl_b200.h>
 *     namespace pcl = pcl_b200;            // or -DPCL_B200_AS_PCL
 *
 * Each adapter flattens its clouds (no copy: the C ABI takes a row stride) and calls the C ABI; the
 * GPU does the work, there is no CPU implementation behind these classes.  Error behaviour follows
 * PCL: a failed initCompute logs to stderr and empties the output, searches return
 * the neighbour count (0 = failure), degenerate points give NaN rows, nothing throws.
 *
 * PCL/Eigen/Boost are not dependencies: std::shared_ptr replaces boost::shared_ptr and Matrix4f is a
 * minimal row-major 4x4 (when this header is used inside a real PCL build, only the algorithm classes
 * are taken from here and the point/container types come from PCL — see INTEGRATION.md).
 */
#ifndef PCL_B200_H_
#define PCL_B200_H_

#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <memory>
#include <mutex>
#include <vector>

#include "../b200reg.h"

namespace pcl_b200 {

/* ---------------------------------------------------------------- point types */
struct alignas(16) PointXYZ {
  float x, y, z, pad;
  PointXYZ() : x(0), y(0), z(0), pad(1.0f) {}
};
struct alignas(16) PointXYZRGBA {  /* SHOT_demo.cpp:38; 32 bytes, colour never used on the hot path */
  float x, y, z, pad;
  uint32_t rgba;
  float pad2[3];
  PointXYZRGBA() : x(0), y(0), z(0), pad(1.0f), rgba(0) { pad2[0] = pad2[1] = pad2[2] = 0; }
};
struct alignas(16) Normal {  /* SHOT_demo.cpp:39 */
  float normal_x, normal_y, normal_z, pad;
  float curvature;
  float pad2[3];
  Normal() : normal_x(0), normal_y(0), normal_z(0), pad(0), curvature(0) { pad2[0] = pad2[1] = pad2[2] = 0; }
};
struct SHOT352 {  /* SHOT_demo.cpp:41 */
  float descriptor[352];
  float rf[9];
  static int descriptorSize() { return 352; }
};
struct FPFHSignature33 {  /* FPFH_demo.cpp:43 */
  float histogram[33];
  static int descriptorSize() { return 33; }
};
struct ReferenceFrame {  /* SHOT_demo.cpp:40 */
  float x_axis[3], y_axis[3], z_axis[3];
};
struct Correspondence {  /* SHOT.cpp:420: (model index, scene index, squared descriptor distance) */
  int index_query;
  int index_match;
  float distance;
  Correspondence() : index_query(0), index_match(-1), distance(3.4028235e38f) {}
  Correspondence(int q, int m, float d) : index_query(q), index_match(m), distance(d) {}
};
static_assert(sizeof(Correspondence) == sizeof(b200_corr), "Correspondence must match b200_corr");
typedef std::vector<Correspondence> Correspondences;
typedef std::shared_ptr<Correspondences> CorrespondencesPtr;
typedef std::shared_ptr<const Correspondences> CorrespondencesConstPtr;

/* row-major 4x4, stands in for Eigen::Matrix4f in recognize()'s output (SHOT.cpp:430) */
struct Matrix4f {
  float m[16];
  Matrix4f() {
    memset(m, 0, sizeof(m));
    m[0] = m[5] = m[10] = m[15] = 1.0f;
  }
  float &operator()(int r, int c) { return m[r * 4 + c]; }
  float operator()(int r, int c) const { return m[r * 4 + c]; }
};

namespace detail {
/* Live clouds of one point type.  The reference's correspondence loop hands KdTreeFLANN::nearestKSearch one element
 * of a descriptor cloud at a time (SHOT.cpp:409-423); the adapter looks the element's address up here to learn that
 * it is row i of a cloud and may answer rows i .. i + B in one device call (see KdTreeFLANN<PointT, true>). */
template <class PointT>
struct CloudRegistry {
  static std::mutex &mu() {
    static std::mutex m;
    return m;
  }
  static std::vector<const std::vector<PointT> *> &live() {
    static std::vector<const std::vector<PointT> *> v;
    return v;
  }
  static void add(const std::vector<PointT> *p) {
    std::lock_guard<std::mutex> lk(mu());
    live().push_back(p);
  }
  static void remove(const std::vector<PointT> *p) {
    std::lock_guard<std::mutex> lk(mu());
    auto &v = live();
    for (size_t i = 0; i < v.size(); ++i)
      if (v[i] == p) {
        v[i] = v.back();
        v.pop_back();
        return;
      }
  }
  /* the live cloud whose storage holds *q (nullptr if none); *index = its row */
  static const std::vector<PointT> *find(const PointT *q, size_t *index) {
    std::lock_guard<std::mutex> lk(mu());
    for (const std::vector<PointT> *v : live()) {
      if (v->empty()) continue;
      const PointT *b = v->data();
      if (q >= b && q < b + v->size()) {
        *index = (size_t)(q - b);
        return v;
      }
    }
    return nullptr;
  }
};
}  // namespace detail

template <class PointT>
struct PointCloud {
  typedef std::shared_ptr<PointCloud<PointT>> Ptr;
  typedef std::shared_ptr<const PointCloud<PointT>> ConstPtr;
  std::vector<PointT> points;
  uint32_t width, height;
  bool is_dense;
  PointCloud() : width(0), height(0), is_dense(true) { detail::CloudRegistry<PointT>::add(&points); }
  PointCloud(const PointCloud &o) : points(o.points), width(o.width), height(o.height), is_dense(o.is_dense) {
    detail::CloudRegistry<PointT>::add(&points);
  }
  PointCloud &operator=(const PointCloud &o) {
    points = o.points;
    width = o.width;
    height = o.height;
    is_dense = o.is_dense;
    return *this;
  }
  ~PointCloud() { detail::CloudRegistry<PointT>::remove(&points); }
  size_t size() const { return points.size(); }
  bool empty() const { return points.empty(); }
  void resize(size_t n) {
    points.resize(n);
    width = (uint32_t)n;
    height = 1;
  }
  void clear() {
    points.clear();
    width = height = 0;
  }
  void push_back(const PointT &p) {
    points.push_back(p);
    width = (uint32_t)points.size();
    height = 1;
  }
  PointT &operator[](size_t i) { return points[i]; }
  const PointT &operator[](size_t i) const { return points[i]; }
  PointT &at(size_t i) { return points.at(i); }
  const PointT &at(size_t i) const { return points.at(i); }
};

namespace detail {

/* one b200_ctx per host thread (SURVEY §8(b) threading: the callers are single-threaded) */
inline b200_ctx *ctx() {
  struct Holder {
    b200_ctx *c = nullptr;
    bool tried = false;
    ~Holder() {
      if (c) b200_ctx_destroy(c);
    }
  };
  static thread_local Holder h;
  if (!h.tried) {
    h.tried = true;
    if (b200_ctx_create(&h.c, 0, nullptr) != B200_OK) {
      fprintf(stderr, "[pcl_b200] no usable B200 device: %s (there is no CPU fallback)\n", b200_last_error(nullptr));
      h.c = nullptr;
    }
  }
  return h.c;
}

inline bool ok(int rc, const char *what) {
  if (rc == B200_OK) return true;
  fprintf(stderr, "[pcl_b200::%s] %s\n", what, b200_last_error(ctx()));
  return false;
}

/* xyz view of a point type: pointer to x and row stride in floats */
template <class PointT>
inline const float *xyz(const std::vector<PointT> &v) {
  return v.empty() ? nullptr : &v[0].x;
}
template <class PointT>
constexpr int stride() {
  return (int)(sizeof(PointT) / sizeof(float));
}

/* descriptor view: pointer, dimension and row stride for KdTreeFLANN<Descriptor> */
template <class T>
struct DescTraits {
  static const bool is_descriptor = false;
};
template <>
struct DescTraits<SHOT352> {
  static const bool is_descriptor = true;
  static const int dim = 352;
  static const float *data(const SHOT352 &p) { return p.descriptor; }
};
template <>
struct DescTraits<FPFHSignature33> {
  static const bool is_descriptor = true;
  static const int dim = 33;
  static const float *data(const FPFHSignature33 &p) { return p.histogram; }
};

/* RAII for a device-resident search surface */
struct Surface {
  b200_cloud *c = nullptr;
  const void *key = nullptr;
  size_t n = 0;
  ~Surface() { reset(); }
  void reset() {
    if (c) b200_cloud_destroy(c);
    c = nullptr;
    key = nullptr;
  }
  template <class PointT>
  bool bind(const std::shared_ptr<const PointCloud<PointT>> &cloud, const char *who) {
    if (!cloud || cloud->empty()) {
      fprintf(stderr, "[pcl_b200::%s] input cloud is empty\n", who);
      reset();
      return false;
    }
    if (c && key == cloud.get() && n == cloud->size()) return true; /* same cloud as before: keep the grids */
    reset();
    if (!ctx()) return false;
    if (!ok(b200_cloud_create(ctx(), xyz(cloud->points), (int)cloud->size(), stride<PointT>(), &c), who)) return false;
    key = cloud.get();
    n = cloud->size();
    return true;
  }
};

}  // namespace detail

/* ---------------------------------------------------------------- search */
/* pcl::KdTreeFLANN<PointT>: xyz points (SHOT_VAR.cpp:350-356, Edge_detection.cpp:117-120) or
 * descriptors (SHOT.cpp:405-417, SHOT_demo.cpp:508-521, FPFH_demo.cpp:516-529). */
template <class PointT, bool IsDesc = detail::DescTraits<PointT>::is_descriptor>
class KdTreeFLANN;

template <class PointT>
class KdTreeFLANN<PointT, false> {
 public:
  typedef std::shared_ptr<KdTreeFLANN<PointT>> Ptr;
  typedef std::shared_ptr<const PointCloud<PointT>> PointCloudConstPtr;
  void setInputCloud(const PointCloudConstPtr &cloud) {
    input_ = cloud;
    surf_.bind(cloud, "KdTreeFLANN::setInputCloud");
  }
  PointCloudConstPtr getInputCloud() const { return input_; }
  int nearestKSearch(const PointT &p, int k, std::vector<int> &k_indices, std::vector<float> &k_sqr_distances) const {
    if (!surf_.c || k < 1) return 0;
    k_indices.assign((size_t)k, -1);
    k_sqr_distances.assign((size_t)k, 0.f);
    int found = 0;
    if (!detail::ok(b200_knn_search(detail::ctx(), surf_.c, &p.x, 1, detail::stride<PointT>(), k, k_indices.data(),
                                    k_sqr_distances.data(), &found),
                    "KdTreeFLANN::nearestKSearch"))
      return 0;
    k_indices.resize((size_t)found);
    k_sqr_distances.resize((size_t)found);
    return found;
  }
  int radiusSearch(const PointT &p, double radius, std::vector<int> &k_indices, std::vector<float> &k_sqr_distances,
                   unsigned int max_nn = 0) const {
    if (!surf_.c) return 0;
    int64_t off[2] = {0, 0}, total = 0;
    if (!detail::ok(b200_radius_search(detail::ctx(), surf_.c, &p.x, 1, detail::stride<PointT>(), radius, off, nullptr,
                                       nullptr, 0, &total),
                    "KdTreeFLANN::radiusSearch"))
      return 0;
    k_indices.assign((size_t)total, 0);
    k_sqr_distances.assign((size_t)total, 0.f);
    if (total > 0 &&
        !detail::ok(b200_radius_search(detail::ctx(), surf_.c, &p.x, 1, detail::stride<PointT>(), radius, off,
                                       k_indices.data(), k_sqr_distances.data(), total, &total),
                    "KdTreeFLANN::radiusSearch"))
      return 0;
    if (max_nn != 0 && (size_t)max_nn < k_indices.size()) { /* sorted ascending: keep the closest */
      k_indices.resize(max_nn);
      k_sqr_distances.resize(max_nn);
    }
    return (int)k_indices.size();
  }
  b200_cloud *handle() const { return surf_.c; }

 private:
  PointCloudConstPtr input_;
  mutable detail::Surface surf_;
};

template <class PointT>
class KdTreeFLANN<PointT, true> {
 public:
  typedef std::shared_ptr<KdTreeFLANN<PointT>> Ptr;
  typedef std::shared_ptr<const PointCloud<PointT>> PointCloudConstPtr;
  ~KdTreeFLANN() { reset(); }
  void setInputCloud(const PointCloudConstPtr &cloud) {
    reset();
    input_ = cloud;
    if (!cloud || cloud->empty()) {
      fprintf(stderr, "[pcl_b200::KdTreeFLANN::setInputCloud] Cannot create a KDTree with an empty input cloud!\n");
      return;
    }
    const int D = detail::DescTraits<PointT>::dim;
    std::vector<float> flat(cloud->size() * (size_t)D);
    for (size_t i = 0; i < cloud->size(); ++i)
      memcpy(&flat[i * D], detail::DescTraits<PointT>::data(cloud->points[i]), sizeof(float) * D);
    if (detail::ctx())
      detail::ok(b200_desc_index_create(detail::ctx(), flat.data(), (int)cloud->size(), D, &index_),
                 "KdTreeFLANN::setInputCloud");
  }
  /* The reference calls this once per scene descriptor from a plain loop (SHOT.cpp:409-423, SHOT_demo.cpp:513-530).
   * One device round trip per call would make that loop ~1000 times slower than the batched search, so when the query
   * is row i of a live descriptor cloud the rows i .. i + lookAhead() are answered in one call and the following calls
   * are served from that answer — after checking, byte for byte, that the row still holds what was sent.  The
   * results are those of the single-query call.  setLookAhead(0) turns it off. */
  int nearestKSearch(const PointT &p, int k, std::vector<int> &k_indices, std::vector<float> &k_sqr_distances) const {
    if (!index_ || k < 1) return 0;
    const int D = detail::DescTraits<PointT>::dim;
    const float *row = detail::DescTraits<PointT>::data(p);
    if (look_ahead_ > 0) {
      size_t i = 0;
      const std::vector<PointT> *cloud = detail::CloudRegistry<PointT>::find(&p, &i);
      if (cloud) {
        const bool hit = cloud == batch_cloud_ && k == batch_k_ && i >= batch_first_ && i < batch_first_ + batch_n_ &&
                         memcmp(row, &batch_q_[(i - batch_first_) * (size_t)D], sizeof(float) * D) == 0;
        if (!hit) {
          const size_t n = std::min(cloud->size() - i, (size_t)look_ahead_);
          batch_q_.resize(n * (size_t)D);
          for (size_t r = 0; r < n; ++r)
            memcpy(&batch_q_[r * D], detail::DescTraits<PointT>::data((*cloud)[i + r]), sizeof(float) * D);
          batch_idx_.assign(n * (size_t)k, -1);
          batch_d2_.assign(n * (size_t)k, 0.f);
          batch_found_ = 0;
          batch_n_ = 0;
          if (!detail::ok(b200_desc_index_knn(detail::ctx(), index_, batch_q_.data(), (int)n, k, batch_idx_.data(),
                                              batch_d2_.data(), &batch_found_),
                          "KdTreeFLANN::nearestKSearch"))
            return 0;
          batch_cloud_ = cloud;
          batch_first_ = i;
          batch_n_ = n;
          batch_k_ = k;
        }
        const size_t o = (i - batch_first_) * (size_t)k;
        int found = 0;
        while (found < batch_found_ && batch_idx_[o + found] >= 0) ++found;
        k_indices.assign(batch_idx_.begin() + o, batch_idx_.begin() + o + found);
        k_sqr_distances.assign(batch_d2_.begin() + o, batch_d2_.begin() + o + found);
        return found;
      }
    }
    k_indices.assign((size_t)k, -1);
    k_sqr_distances.assign((size_t)k, 0.f);
    int found = 0;
    if (!detail::ok(b200_desc_index_knn(detail::ctx(), index_, row, 1, k, k_indices.data(), k_sqr_distances.data(),
                                        &found),
                    "KdTreeFLANN::nearestKSearch"))
      return 0;
    k_indices.resize((size_t)found);
    k_sqr_distances.resize((size_t)found);
    return found;
  }
  void setLookAhead(int rows) { look_ahead_ = rows < 0 ? 0 : rows; }
  int lookAhead() const { return look_ahead_; }

 private:
  void reset() {
    if (index_) b200_desc_index_destroy(index_);
    index_ = nullptr;
    batch_cloud_ = nullptr;
    batch_n_ = 0;
  }
  PointCloudConstPtr input_;
  b200_desc_index *index_ = nullptr;
  int look_ahead_ = 8192;
  mutable const std::vector<PointT> *batch_cloud_ = nullptr;
  mutable size_t batch_first_ = 0, batch_n_ = 0;
  mutable int batch_k_ = 0, batch_found_ = 0;
  mutable std::vector<float> batch_q_, batch_d2_;
  mutable std::vector<int> batch_idx_;
};

namespace search {
/* pcl::search::KdTree<PointT> (SHOT_demo.cpp:294, 404): same search calls; features accept it through
 * setSearchMethod and keep using their own device-resident surface. */
template <class PointT>
class KdTree : public KdTreeFLANN<PointT> {
 public:
  typedef std::shared_ptr<KdTree<PointT>> Ptr;
};
}  // namespace search

/* Batched form of the reference's correspondence loop (SHOT.cpp:403-424 mode 1, SHOT_demo.cpp:508-531
 * mode 2): one call instead of K_s nearestKSearch calls.  Same output, ascending scene index. */
template <class DescT>
inline bool determineCorrespondences(const PointCloud<DescT> &model, const PointCloud<DescT> &scene, int mode, float thr,
                                     Correspondences &out) {
  out.clear();
  if (model.empty() || scene.empty() || !detail::ctx()) return false;
  const int D = detail::DescTraits<DescT>::dim;
  std::vector<float> fm(model.size() * (size_t)D), fs(scene.size() * (size_t)D);
  for (size_t i = 0; i < model.size(); ++i) memcpy(&fm[i * D], detail::DescTraits<DescT>::data(model.points[i]), 4 * D);
  for (size_t i = 0; i < scene.size(); ++i) memcpy(&fs[i * D], detail::DescTraits<DescT>::data(scene.points[i]), 4 * D);
  out.resize(scene.size());
  int count = 0;
  const bool good = detail::ok(b200_match(detail::ctx(), fm.data(), (int)model.size(), fs.data(), (int)scene.size(), D, mode,
                                          thr, reinterpret_cast<b200_corr *>(out.data()), &count),
                               "determineCorrespondences");
  out.resize(good ? (size_t)count : 0);
  return good;
}

/* ---------------------------------------------------------------- keypoints */
/* pcl::UniformSampling<PointT> (SHOT.cpp:314-323): setInputCloud, setRadiusSearch(leaf), filter(output).
 * Keeps input points (all fields are copied); output order is ascending leaf index. */
template <class PointT>
class UniformSampling {
 public:
  typedef std::shared_ptr<const PointCloud<PointT>> PointCloudConstPtr;
  void setInputCloud(const PointCloudConstPtr &cloud) { input_ = cloud; }
  void setRadiusSearch(double radius) { leaf_ = radius; }
  void filter(PointCloud<PointT> &output) {
    output.clear();
    if (!input_ || input_->empty() || !detail::ctx()) return;
    const int n = (int)input_->size();
    std::vector<float> xyz((size_t)n * 3);
    std::vector<int> idx((size_t)n);
    int count = 0;
    if (!detail::ok(b200_uniform_sampling(detail::ctx(), detail::xyz(input_->points), n, detail::stride<PointT>(), leaf_,
                                          xyz.data(), idx.data(), &count),
                    "UniformSampling::filter"))
      return;
    output.points.resize((size_t)count);
    for (int i = 0; i < count; ++i) output.points[i] = input_->points[idx[i]];
    output.width = (uint32_t)count;
    output.height = 1;
    output.is_dense = true;
  }

 private:
  PointCloudConstPtr input_;
  double leaf_ = 0.0;
};

/* pcl::VoxelGrid<PointT> (SHOT_demo.cpp:413-417, 489-491): setInputCloud, setLeafSize, filter(output).
 * x, y, z are the voxel centroids; the other fields (colour) are left at their defaults. */
template <class PointT>
class VoxelGrid {
 public:
  typedef std::shared_ptr<const PointCloud<PointT>> PointCloudConstPtr;
  void setInputCloud(const PointCloudConstPtr &cloud) { input_ = cloud; }
  void setLeafSize(float lx, float ly, float lz) { l_[0] = lx, l_[1] = ly, l_[2] = lz; }
  void filter(PointCloud<PointT> &output) {
    output.clear();
    if (!input_ || input_->empty() || !detail::ctx()) return;
    const int n = (int)input_->size();
    std::vector<float> xyz((size_t)n * 3);
    int count = 0;
    if (!detail::ok(b200_voxel_grid(detail::ctx(), detail::xyz(input_->points), n, detail::stride<PointT>(), l_[0], l_[1],
                                    l_[2], xyz.data(), &count),
                    "VoxelGrid::filter"))
      return;
    output.points.resize((size_t)count);
    for (int i = 0; i < count; ++i) {
      output.points[i].x = xyz[(size_t)i * 3], output.points[i].y = xyz[(size_t)i * 3 + 1];
      output.points[i].z = xyz[(size_t)i * 3 + 2];
    }
    output.width = (uint32_t)count;
    output.height = 1;
    output.is_dense = true;
  }

 private:
  PointCloudConstPtr input_;
  float l_[3] = {0.f, 0.f, 0.f};
};

/* ---------------------------------------------------------------- Feature base */
template <class PointInT>
class FeatureBase {
 public:
  typedef std::shared_ptr<const PointCloud<PointInT>> PointCloudInConstPtr;
  void setInputCloud(const PointCloudInConstPtr &cloud) { input_ = cloud; }
  void setSearchSurface(const PointCloudInConstPtr &cloud) { surface_ = cloud; }
  void setKSearch(int k) { k_ = k; }
  void setRadiusSearch(double r) { search_radius_ = r; }
  template <class TreePtr>
  void setSearchMethod(const TreePtr &) {} /* the device grid replaces the kd-tree */
  void setNumberOfThreads(unsigned int) {}
  int getKSearch() const { return k_; }
  double getRadiusSearch() const { return search_radius_; }

 protected:
  /* Feature::initCompute: surface defaults to the input; exactly one of k / radius */
  bool initCompute(const char *who, bool radius_only) {
    if (!input_ || input_->empty()) {
      fprintf(stderr, "[pcl_b200::%s::initCompute] input cloud is empty\n", who);
      return false;
    }
    if (!surface_) fake_surface_ = true;
    const PointCloudInConstPtr &surf = fake_surface_ ? input_ : surface_;
    if (search_radius_ != 0.0 && k_ != 0) {
      fprintf(stderr, "[pcl_b200::%s::initCompute] Both radius (%f) and K (%d) defined! Set one of them to zero first.\n",
              who, search_radius_, k_);
      return false;
    }
    if (search_radius_ == 0.0 && (k_ == 0 || radius_only)) {
      fprintf(stderr, "[pcl_b200::%s::initCompute] Neither radius nor K defined!\n", who);
      return false;
    }
    return surf_.bind(surf, who);
  }
  void deinitCompute() {
    if (fake_surface_) {
      fake_surface_ = false;
      surface_.reset();
    }
  }
  bool inputIsSurface() const { return fake_surface_ || surface_.get() == input_.get(); }
  PointCloudInConstPtr input_, surface_;
  bool fake_surface_ = false;
  int k_ = 0;
  double search_radius_ = 0.0;
  detail::Surface surf_;
};

/* ---------------------------------------------------------------- cloud utilities */
/* pcl::removeNaNFromPointCloud (SHOT.cpp:298-299): keeps the rows with finite x, y, z (all fields of the kept
 * points are copied), index = their positions in the input; cloud_out becomes dense and unorganised. */
template <class PointT>
inline void removeNaNFromPointCloud(const PointCloud<PointT> &cloud_in, PointCloud<PointT> &cloud_out,
                                    std::vector<int> &index) {
  const size_t n = cloud_in.size();
  std::vector<int> idx(n ? n : 1);
  std::vector<float> xyz((n ? n : 1) * 3);
  int kept = 0;
  if (n && (!detail::ctx() ||
            !detail::ok(b200_remove_nan(detail::ctx(), detail::xyz(cloud_in.points), (int)n, detail::stride<PointT>(),
                                        xyz.data(), idx.data(), &kept),
                        "removeNaNFromPointCloud")))
    kept = 0;
  PointCloud<PointT> out;
  out.points.resize((size_t)kept);
  for (int i = 0; i < kept; ++i) out.points[(size_t)i] = cloud_in.points[(size_t)idx[(size_t)i]];
  out.width = (uint32_t)kept;
  out.height = 1;
  out.is_dense = true;
  index.assign(idx.begin(), idx.begin() + kept);
  cloud_out = out; /* in-place use (cloud_in == cloud_out) is what the reference does */
}

/* pcl::transformPointCloud(cloud_in, cloud_out, transform) with a 4x4 matrix (the model placed by a pose before
 * ICP and for display, SHOT_demo.cpp:590-663). */
template <class PointT>
inline void transformPointCloud(const PointCloud<PointT> &cloud_in, PointCloud<PointT> &cloud_out, const Matrix4f &transform) {
  const size_t n = cloud_in.size();
  std::vector<float> xyz((n ? n : 1) * 3);
  if (n && (!detail::ctx() ||
            !detail::ok(b200_transform_points(detail::ctx(), detail::xyz(cloud_in.points), (int)n, detail::stride<PointT>(),
                                              transform.m, xyz.data()),
                        "transformPointCloud")))
    return;
  PointCloud<PointT> out = cloud_in;
  for (size_t i = 0; i < n; ++i) {
    out.points[i].x = xyz[3 * i + 0];
    out.points[i].y = xyz[3 * i + 1];
    out.points[i].z = xyz[3 * i + 2];
  }
  cloud_out = out;
}

/* ---------------------------------------------------------------- normals (a10) */
/* pcl::NormalEstimationOMP (SHOT.cpp:302-308, SHOT_demo.cpp:405-411, FPFH_demo.cpp:416-420) */
template <class PointInT, class PointOutT = Normal>
class NormalEstimationOMP : public FeatureBase<PointInT> {
 public:
  void setViewPoint(float vx, float vy, float vz) {
    vp_[0] = vx, vp_[1] = vy, vp_[2] = vz;
  }
  void compute(PointCloud<PointOutT> &output) {
    output.clear();
    if (!this->initCompute("NormalEstimationOMP", false)) return;
    const size_t n = this->input_->size();
    std::vector<float> out(n * 4);
    const bool self = this->inputIsSurface();
    const int rc = b200_normals(detail::ctx(), this->surf_.c, self ? nullptr : detail::xyz(this->input_->points), (int)n,
                                detail::stride<PointInT>(), this->k_, this->search_radius_, vp_, out.data());
    this->deinitCompute();
    if (!detail::ok(rc, "NormalEstimationOMP::compute")) return;
    output.points.resize(n);
    output.width = this->input_->width ? this->input_->width : (uint32_t)n;
    output.height = this->input_->height ? this->input_->height : 1;
    output.is_dense = true;
    for (size_t i = 0; i < n; ++i) {
      PointOutT &p = output.points[i];
      p.normal_x = out[i * 4 + 0], p.normal_y = out[i * 4 + 1], p.normal_z = out[i * 4 + 2];
      p.curvature = out[i * 4 + 3];
      if (p.normal_x != p.normal_x) output.is_dense = false;
    }
  }

 private:
  float vp_[3] = {0.f, 0.f, 0.f};
};
template <class PointInT, class PointOutT = Normal>
using NormalEstimation = NormalEstimationOMP<PointInT, PointOutT>;

namespace detail {
template <class PointNT>
inline std::vector<float> flatten_normals(const PointCloud<PointNT> &nrm) {
  std::vector<float> f(nrm.size() * 4);
  for (size_t i = 0; i < nrm.size(); ++i) {
    f[i * 4 + 0] = nrm.points[i].normal_x, f[i * 4 + 1] = nrm.points[i].normal_y, f[i * 4 + 2] = nrm.points[i].normal_z;
    f[i * 4 + 3] = nrm.points[i].curvature;
  }
  return f;
}
}  // namespace detail

/* ---------------------------------------------------------------- SHOT (a11, a12) */
/* pcl::SHOTEstimationOMP<PointXYZRGBA, Normal, SHOT352> (SHOT.cpp:360-371, SHOT_demo.cpp:419-424,
 * 497-502, CAD_desc.cpp:341-352).  The local reference frames are computed inside, as PCL does. */
template <class PointInT, class PointNT = Normal, class PointOutT = SHOT352>
class SHOTEstimationOMP : public FeatureBase<PointInT> {
 public:
  typedef std::shared_ptr<const PointCloud<PointNT>> PointCloudNConstPtr;
  void setInputNormals(const PointCloudNConstPtr &normals) { normals_ = normals; }
  void compute(PointCloud<PointOutT> &output) {
    output.clear();
    if (!this->initCompute("SHOTEstimationOMP", true)) return;
    const size_t n_surf = (this->fake_surface_ ? this->input_ : this->surface_)->size();
    if (!normals_ || normals_->size() != n_surf) {
      fprintf(stderr,
              "[pcl_b200::SHOTEstimationOMP::initCompute] The number of points in the surface differs from the number "
              "of normals!\n");
      this->deinitCompute();
      return;
    }
    const size_t K = this->input_->size();
    std::vector<float> nrm = detail::flatten_normals(*normals_), desc(K * 352), rf(K * 9);
    const int rc = b200_shot352(detail::ctx(), this->surf_.c, nrm.data(), detail::xyz(this->input_->points), (int)K,
                                detail::stride<PointInT>(), this->search_radius_, desc.data(), rf.data());
    this->deinitCompute();
    if (!detail::ok(rc, "SHOTEstimationOMP::compute")) return;
    output.points.resize(K);
    output.width = (uint32_t)K;
    output.height = 1;
    output.is_dense = true;
    for (size_t i = 0; i < K; ++i) {
      memcpy(output.points[i].descriptor, &desc[i * 352], sizeof(float) * 352);
      memcpy(output.points[i].rf, &rf[i * 9], sizeof(float) * 9);
      if (desc[i * 352] != desc[i * 352]) output.is_dense = false;
    }
  }

 private:
  PointCloudNConstPtr normals_;
};
template <class PointInT, class PointNT = Normal, class PointOutT = SHOT352>
using SHOTEstimation = SHOTEstimationOMP<PointInT, PointNT, PointOutT>;

/* ---------------------------------------------------------------- BOARD frames */
/* pcl::BOARDLocalReferenceFrameEstimation<PointInT, PointNT, ReferenceFrame> (SHOT.cpp:441-453, 6Dpose.cpp:497-509,
 * FPFH_demo.cpp:556-568).  PCL draws the random reference axis from rand(); the device library keeps its own
 * glibc-compatible stream per context (b200_ctx_srand; a fresh context is srand(1)). */
template <class PointInT, class PointNT = Normal, class PointOutT = ReferenceFrame>
class BOARDLocalReferenceFrameEstimation : public FeatureBase<PointInT> {
 public:
  typedef std::shared_ptr<const PointCloud<PointNT>> PointCloudNConstPtr;
  BOARDLocalReferenceFrameEstimation() { b200_board_params_default(&params_); }
  void setInputNormals(const PointCloudNConstPtr &normals) { normals_ = normals; }
  void setFindHoles(bool find_holes) { params_.find_holes = find_holes ? 1 : 0; }
  bool getFindHoles() const { return params_.find_holes != 0; }
  void setTangentRadius(float radius) { params_.tangent_radius = radius; }
  void setMarginThresh(float margin_thresh) { params_.margin_thresh = margin_thresh; }
  void setCheckMarginArraySize(int size) { params_.check_margin_array_size = size; }
  void setHoleSizeProbThresh(float prob_thresh) { params_.hole_size_prob_thresh = prob_thresh; }
  void setSteepThresh(float steep_thresh) { params_.steep_thresh = steep_thresh; }
  void compute(PointCloud<PointOutT> &output) {
    output.clear();
    if (this->k_ != 0) {
      fprintf(stderr,
              "[pcl_b200::BOARDLocalReferenceFrameEstimation::computeFeature] Error! Search method set to k-neighborhood. "
              "Call setKSearch(0) and setRadiusSearch( radius ) to use this class.\n");
      return;
    }
    if (!this->initCompute("BOARDLocalReferenceFrameEstimation", true)) return;
    const size_t n_surf = (this->fake_surface_ ? this->input_ : this->surface_)->size();
    if (!normals_ || normals_->size() != n_surf) {
      fprintf(stderr,
              "[pcl_b200::BOARDLocalReferenceFrameEstimation::initCompute] The number of points in the surface differs "
              "from the number of normals!\n");
      this->deinitCompute();
      return;
    }
    const size_t K = this->input_->size();
    std::vector<float> nrm = detail::flatten_normals(*normals_), rf(K * 9);
    const int rc = b200_board_lrf(detail::ctx(), this->surf_.c, nrm.data(), detail::xyz(this->input_->points), (int)K,
                                  detail::stride<PointInT>(), this->search_radius_, &params_, rf.data());
    this->deinitCompute();
    if (!detail::ok(rc, "BOARDLocalReferenceFrameEstimation::compute")) return;
    output.points.resize(K);
    output.width = (uint32_t)K;
    output.height = 1;
    output.is_dense = true;
    for (size_t i = 0; i < K; ++i) {
      memcpy(output.points[i].x_axis, &rf[i * 9 + 0], 12);
      memcpy(output.points[i].y_axis, &rf[i * 9 + 3], 12);
      memcpy(output.points[i].z_axis, &rf[i * 9 + 6], 12);
      if (rf[i * 9] != rf[i * 9]) output.is_dense = false;
    }
  }

 private:
  PointCloudNConstPtr normals_;
  b200_board_params params_;
};

/* ---------------------------------------------------------------- FPFH (a13) */
/* pcl::FPFHEstimation / FPFHEstimationOMP (FPFH_demo.cpp:422-428, 505-510;
 * FPFH_scenes_clustered.cpp:287-293, 379-387) */
template <class PointInT, class PointNT = Normal, class PointOutT = FPFHSignature33>
class FPFHEstimationOMP : public FeatureBase<PointInT> {
 public:
  typedef std::shared_ptr<const PointCloud<PointNT>> PointCloudNConstPtr;
  void setInputNormals(const PointCloudNConstPtr &normals) { normals_ = normals; }
  void compute(PointCloud<PointOutT> &output) {
    output.clear();
    if (!this->initCompute("FPFHEstimation", true)) return;
    const size_t n_surf = (this->fake_surface_ ? this->input_ : this->surface_)->size();
    if (!normals_ || normals_->size() != n_surf) {
      fprintf(stderr,
              "[pcl_b200::FPFHEstimation::initCompute] The number of points in the surface differs from the number of "
              "normals!\n");
      this->deinitCompute();
      return;
    }
    const size_t K = this->input_->size();
    std::vector<float> nrm = detail::flatten_normals(*normals_), out(K * 33);
    const bool self = this->inputIsSurface();
    const int rc = b200_fpfh33(detail::ctx(), this->surf_.c, nrm.data(), self ? nullptr : detail::xyz(this->input_->points),
                               (int)K, detail::stride<PointInT>(), this->search_radius_, out.data());
    this->deinitCompute();
    if (!detail::ok(rc, "FPFHEstimation::compute")) return;
    output.points.resize(K);
    output.width = (uint32_t)K;
    output.height = 1;
    output.is_dense = true;
    for (size_t i = 0; i < K; ++i) {
      memcpy(output.points[i].histogram, &out[i * 33], sizeof(float) * 33);
      if (out[i * 33] != out[i * 33]) output.is_dense = false;
    }
  }

 private:
  PointCloudNConstPtr normals_;
};
template <class PointInT, class PointNT = Normal, class PointOutT = FPFHSignature33>
using FPFHEstimation = FPFHEstimationOMP<PointInT, PointNT, PointOutT>;

/* ---------------------------------------------------------------- grouping (a15) */
/* pcl::GeometricConsistencyGrouping<PointXYZRGBA, PointXYZRGBA> (SHOT.cpp:473-482, 6Dpose.cpp:529-538) */
template <class PointModelT, class PointSceneT>
class GeometricConsistencyGrouping {
 public:
  typedef std::shared_ptr<const PointCloud<PointModelT>> PointCloudConstPtr;
  typedef std::shared_ptr<const PointCloud<PointSceneT>> SceneCloudConstPtr;
  void setGCSize(double gc_size) { gc_size_ = gc_size; }
  void setGCThreshold(int threshold) { gc_threshold_ = threshold; } /* the reference passes 2.0f / 3.0f: truncated */
  double getGCSize() const { return gc_size_; }
  int getGCThreshold() const { return gc_threshold_; }
  void setInputCloud(const PointCloudConstPtr &cloud) { input_ = cloud; }
  void setSceneCloud(const SceneCloudConstPtr &scene) { scene_ = scene; }
  void setModelSceneCorrespondences(const CorrespondencesConstPtr &corrs) { model_scene_corrs_ = corrs; }
  bool recognize(std::vector<Matrix4f> &transformations) {
    std::vector<Correspondences> clustered;
    return recognize(transformations, clustered);
  }
  bool recognize(std::vector<Matrix4f> &transformations, std::vector<Correspondences> &clustered_corrs) {
    transformations.clear();
    clustered_corrs.clear();
    if (!input_ || !scene_ || input_->empty() || scene_->empty()) {
      fprintf(stderr, "[pcl_b200::GeometricConsistencyGrouping::recognize] model or scene cloud not set\n");
      return false;
    }
    if (!model_scene_corrs_ || model_scene_corrs_->empty()) {
      fprintf(stderr,
              "[pcl_b200::GeometricConsistencyGrouping::clusterCorrespondences()] Error! Correspondences not set, please "
              "set them before calling again this function.\n");
      return false;
    }
    if (!detail::ctx()) return false;
    const int C = (int)model_scene_corrs_->size();
    int max_inst = C / (gc_threshold_ > 0 ? gc_threshold_ + 1 : 1) + 1; /* a set has more than gc_threshold members */
    std::vector<float> T((size_t)max_inst * 16);
    std::vector<int> off((size_t)max_inst + 1);
    std::vector<b200_corr> out((size_t)C);
    int n = 0;
    const int rc = b200_gc_recognize(detail::ctx(), detail::xyz(input_->points), (int)input_->size(),
                                     detail::stride<PointModelT>(), detail::xyz(scene_->points), (int)scene_->size(),
                                     detail::stride<PointSceneT>(),
                                     reinterpret_cast<const b200_corr *>(model_scene_corrs_->data()), C, gc_size_,
                                     gc_threshold_, T.data(), max_inst, off.data(), out.data(), C, &n);
    if (!detail::ok(rc, "GeometricConsistencyGrouping::recognize")) return false;
    transformations.resize((size_t)n);
    clustered_corrs.resize((size_t)n);
    for (int i = 0; i < n; ++i) {
      memcpy(transformations[i].m, &T[(size_t)i * 16], sizeof(float) * 16);
      clustered_corrs[i].resize((size_t)(off[i + 1] - off[i]));
      if (off[i + 1] > off[i])
        memcpy(static_cast<void *>(clustered_corrs[i].data()), &out[off[i]], sizeof(b200_corr) * (size_t)(off[i + 1] - off[i]));
    }
    return true;
  }

 private:
  PointCloudConstPtr input_;
  SceneCloudConstPtr scene_;
  CorrespondencesConstPtr model_scene_corrs_;
  double gc_size_ = 1.0;
  int gc_threshold_ = 3;
};

/* pcl::Hough3DGrouping<PointModelT, PointSceneT, ReferenceFrame, ReferenceFrame> (SHOT.cpp:433-470,
 * SHOT_demo.cpp:540-577): the reference passes the frames with setInputRf / setSceneRf and uses
 * setUseInterpolation(false), setUseDistanceWeight(true).  Interpolated voting is not provided. */
template <class PointModelT, class PointSceneT, class PointModelRfT = ReferenceFrame, class PointSceneRfT = ReferenceFrame>
class Hough3DGrouping {
 public:
  typedef std::shared_ptr<const PointCloud<PointModelT>> PointCloudConstPtr;
  typedef std::shared_ptr<const PointCloud<PointSceneT>> SceneCloudConstPtr;
  typedef std::shared_ptr<const PointCloud<PointModelRfT>> ModelRfCloudConstPtr;
  typedef std::shared_ptr<const PointCloud<PointSceneRfT>> SceneRfCloudConstPtr;
  void setHoughBinSize(double bin_size) { bin_size_ = bin_size; }
  void setHoughThreshold(double threshold) { threshold_ = threshold; }
  void setUseInterpolation(bool use) { use_interpolation_ = use; }
  void setUseDistanceWeight(bool) {} /* without interpolation PCL's weight is 1 either way (see hough.cu) */
  void setInputCloud(const PointCloudConstPtr &cloud) { input_ = cloud; }
  void setInputRf(const ModelRfCloudConstPtr &rf) { input_rf_ = rf; }
  void setSceneCloud(const SceneCloudConstPtr &scene) { scene_ = scene; }
  void setSceneRf(const SceneRfCloudConstPtr &rf) { scene_rf_ = rf; }
  void setModelSceneCorrespondences(const CorrespondencesConstPtr &corrs) { model_scene_corrs_ = corrs; }
  bool recognize(std::vector<Matrix4f> &transformations) {
    std::vector<Correspondences> clustered;
    return recognize(transformations, clustered);
  }
  bool recognize(std::vector<Matrix4f> &transformations, std::vector<Correspondences> &clustered_corrs) {
    transformations.clear();
    clustered_corrs.clear();
    if (!input_ || !scene_ || input_->empty() || scene_->empty() || !input_rf_ || !scene_rf_ ||
        input_rf_->size() != input_->size() || scene_rf_->size() != scene_->size()) {
      fprintf(stderr, "[pcl_b200::Hough3DGrouping::recognize] clouds and reference frames must be set and of equal size\n");
      return false;
    }
    if (use_interpolation_) {
      fprintf(stderr, "[pcl_b200::Hough3DGrouping::recognize] interpolated voting is not provided\n");
      return false;
    }
    if (!model_scene_corrs_ || model_scene_corrs_->empty()) {
      fprintf(stderr, "[pcl_b200::Hough3DGrouping::recognize] Error! Correspondences not set\n");
      return false;
    }
    if (!detail::ctx()) return false;
    auto flatten = [](const auto &rf_cloud) {
      std::vector<float> f(rf_cloud.size() * 9);
      for (size_t i = 0; i < rf_cloud.size(); ++i) {
        memcpy(&f[i * 9 + 0], rf_cloud.points[i].x_axis, 12);
        memcpy(&f[i * 9 + 3], rf_cloud.points[i].y_axis, 12);
        memcpy(&f[i * 9 + 6], rf_cloud.points[i].z_axis, 12);
      }
      return f;
    };
    const std::vector<float> mrf = flatten(*input_rf_), srf = flatten(*scene_rf_);
    const int C = (int)model_scene_corrs_->size();
    const int max_inst = C;
    std::vector<float> T((size_t)max_inst * 16);
    std::vector<int> off((size_t)max_inst + 1);
    std::vector<b200_corr> out((size_t)C);
    int n = 0;
    const int rc = b200_hough3d_recognize(
        detail::ctx(), detail::xyz(input_->points), mrf.data(), (int)input_->size(), detail::stride<PointModelT>(),
        detail::xyz(scene_->points), srf.data(), (int)scene_->size(), detail::stride<PointSceneT>(),
        reinterpret_cast<const b200_corr *>(model_scene_corrs_->data()), C, bin_size_, threshold_, T.data(), max_inst,
        off.data(), out.data(), C, &n);
    if (!detail::ok(rc, "Hough3DGrouping::recognize")) return false;
    transformations.resize((size_t)n);
    clustered_corrs.resize((size_t)n);
    for (int i = 0; i < n; ++i) {
      memcpy(transformations[i].m, &T[(size_t)i * 16], sizeof(float) * 16);
      clustered_corrs[i].resize((size_t)(off[i + 1] - off[i]));
      if (off[i + 1] > off[i])
        memcpy(static_cast<void *>(clustered_corrs[i].data()), &out[off[i]], sizeof(b200_corr) * (size_t)(off[i + 1] - off[i]));
    }
    return true;
  }

 private:
  PointCloudConstPtr input_;
  SceneCloudConstPtr scene_;
  ModelRfCloudConstPtr input_rf_;
  SceneRfCloudConstPtr scene_rf_;
  CorrespondencesConstPtr model_scene_corrs_;
  double bin_size_ = 1.0, threshold_ = 1.0;
  bool use_interpolation_ = false;
};

/* pcl::IterativeClosestPoint<PointSource, PointTarget> (SHOT.cpp:177-192, SHOT_demo.cpp:604-663,
 * FPFH_demo.cpp:611-668, 6Dpose.cpp:572-609): setMaximumIterations / setInputSource / setInputTarget / align /
 * hasConverged / getFitnessScore / getFinalTransformation, plus the setters the reference leaves at their defaults. */
template <class PointSource, class PointTarget>
class IterativeClosestPoint {
 public:
  typedef std::shared_ptr<const PointCloud<PointSource>> PointCloudSourceConstPtr;
  typedef std::shared_ptr<const PointCloud<PointTarget>> PointCloudTargetConstPtr;
  void setMaximumIterations(int n) { max_iterations_ = n; }
  int getMaximumIterations() const { return max_iterations_; }
  void setMaxCorrespondenceDistance(double d) { corr_dist_threshold_ = d; }
  void setTransformationEpsilon(double e) { transformation_epsilon_ = e; }
  void setEuclideanFitnessEpsilon(double e) { euclidean_fitness_epsilon_ = e; }
  void setInputSource(const PointCloudSourceConstPtr &cloud) { source_ = cloud; }
  void setInputTarget(const PointCloudTargetConstPtr &cloud) {
    target_ = cloud;
    surf_.bind(cloud, "IterativeClosestPoint::setInputTarget");
  }
  void align(PointCloud<PointSource> &output) { align(output, Matrix4f()); }
  void align(PointCloud<PointSource> &output, const Matrix4f &guess) {
    converged_ = false;
    final_transformation_ = guess;
    fitness_ = 1.7976931348623157e308;
    if (!source_ || source_->empty() || !surf_.c) {
      fprintf(stderr, "[pcl_b200::IterativeClosestPoint::align] source or target cloud not set\n");
      return;
    }
    if (!detail::ctx()) return;
    const int ns = (int)source_->size();
    std::vector<float> al((size_t)ns * 3);
    int conv = 0, it = 0;
    if (!detail::ok(b200_icp_align(detail::ctx(), detail::xyz(source_->points), ns, detail::stride<PointSource>(), surf_.c,
                                   max_iterations_, corr_dist_threshold_, transformation_epsilon_,
                                   euclidean_fitness_epsilon_, guess.m, final_transformation_.m, al.data(), &fitness_,
                                   &conv, &it),
                    "IterativeClosestPoint::align"))
      return;
    converged_ = conv != 0;
    output = *source_; /* PCL copies the source's fields and overwrites x, y, z */
    for (int i = 0; i < ns; ++i) {
      output.points[(size_t)i].x = al[3 * (size_t)i + 0];
      output.points[(size_t)i].y = al[3 * (size_t)i + 1];
      output.points[(size_t)i].z = al[3 * (size_t)i + 2];
    }
  }
  bool hasConverged() const { return converged_; }
  double getFitnessScore() const { return fitness_; }
  Matrix4f getFinalTransformation() const { return final_transformation_; }

 private:
  PointCloudSourceConstPtr source_;
  PointCloudTargetConstPtr target_;
  detail::Surface surf_;
  int max_iterations_ = 10;
  double corr_dist_threshold_ = 0.0; /* <= 0: PCL's default sqrt(DBL_MAX), i.e. unlimited */
  double transformation_epsilon_ = 0.0;
  double euclidean_fitness_epsilon_ = -1.7976931348623157e308;
  bool converged_ = false;
  double fitness_ = 1.7976931348623157e308;
  Matrix4f final_transformation_;
};

/* ---- pcl::GlobalHypothesesVerification<ModelT, SceneT> (SHOT_hypothesis.cpp:631-653) --------------------------
 * Same call order, same order dependence as PCL: setSceneCloud voxelises the scene at once, addModels filters the
 * hypotheses with the occlusion threshold in force at that moment, the other setters are read by verify(). */
template <class ModelT, class SceneT>
class GlobalHypothesesVerification {
 public:
  GlobalHypothesesVerification() { b200_hv_params_default(&p_); }
  ~GlobalHypothesesVerification() {
    if (hv_) b200_hv_destroy(hv_);
  }
  GlobalHypothesesVerification(const GlobalHypothesesVerification &) = delete;
  GlobalHypothesesVerification &operator=(const GlobalHypothesesVerification &) = delete;
  void setResolution(float r) { p_.resolution = r, push(); }
  void setInlierThreshold(float r) { p_.inlier_threshold = r, push(); }
  void setOcclusionThreshold(float t) { p_.occlusion_threshold = t, push(); }
  void setRegularizer(float r) { p_.regularizer = r, push(); }
  void setRadiusClutter(float r) { p_.radius_clutter = r, push(); }
  void setClutterRegularizer(float r) { p_.clutter_regularizer = r, push(); }
  void setDetectClutter(bool d) { p_.detect_clutter = d ? 1 : 0, push(); }
  void setRadiusNormals(float r) { p_.radius_normals = r, push(); }
  void setMaxIterations(int i) { p_.max_iterations = i, push(); }
  void setInitialTemp(float t) { p_.initial_temp = t, push(); }
  void setSceneCloud(const typename PointCloud<SceneT>::Ptr &scene) { setSceneCloud(typename PointCloud<SceneT>::ConstPtr(scene)); }
  void setSceneCloud(const typename PointCloud<SceneT>::ConstPtr &scene) {
    n_models_ = 0;
    mask_.clear();
    if (!handle() || !scene) return;
    scene_ok_ = detail::ok(b200_hv_set_scene(detail::ctx(), hv_, detail::xyz(scene->points), (int)scene->size(),
                                             detail::stride<SceneT>()),
                           "GlobalHypothesesVerification::setSceneCloud");
  }
  void addModels(std::vector<typename PointCloud<ModelT>::ConstPtr> &models, bool occlusion_reasoning = false) {
    mask_.clear();
    n_models_ = 0;
    if (!handle()) return;
    std::vector<int> off(1, 0);
    std::vector<float> flat;
    for (const auto &m : models) {
      const size_t n = m ? m->size() : 0;
      for (size_t i = 0; i < n; ++i) {
        flat.push_back(m->points[i].x);
        flat.push_back(m->points[i].y);
        flat.push_back(m->points[i].z);
      }
      off.push_back((int)(flat.size() / 3));
    }
    if (flat.empty()) flat.resize(3, 0.f);
    if (detail::ok(b200_hv_add_models(detail::ctx(), hv_, flat.data(), off.data(), (int)models.size(), 3,
                                      occlusion_reasoning ? 1 : 0),
                   "GlobalHypothesesVerification::addModels"))
      n_models_ = (int)models.size();
  }
  void addModels(std::vector<typename PointCloud<ModelT>::Ptr> &models, bool occlusion_reasoning = false) {
    std::vector<typename PointCloud<ModelT>::ConstPtr> c(models.begin(), models.end());
    addModels(c, occlusion_reasoning);
  }
  void verify() {
    mask_.assign((size_t)n_models_, false);
    if (!handle() || n_models_ == 0) return;
    std::vector<unsigned char> m((size_t)n_models_, 0);
    if (!detail::ok(b200_hv_verify(detail::ctx(), hv_, m.data(), nullptr, &best_cost_, &accepted_moves_),
                    "GlobalHypothesesVerification::verify"))
      return;
    for (int i = 0; i < n_models_; ++i) mask_[(size_t)i] = m[(size_t)i] != 0;
  }
  void getMask(std::vector<bool> &mask) const { mask = mask_; }
  double getBestCost() const { return best_cost_; } /* not in PCL: the annealing's best cost */

 private:
  bool handle() {
    if (hv_) return true;
    if (!detail::ctx()) return false;
    return detail::ok(b200_hv_create(detail::ctx(), &p_, &hv_), "GlobalHypothesesVerification");
  }
  void push() {
    if (hv_) b200_hv_set_params(hv_, &p_);
  }
  b200_hv *hv_ = nullptr;
  b200_hv_params p_;
  int n_models_ = 0;
  bool scene_ok_ = false;
  std::vector<bool> mask_;
  double best_cost_ = 0.0;
  int accepted_moves_ = 0;
};

}  // namespace pcl_b200

#ifdef PCL_B200_AS_PCL
namespace pcl = pcl_b200;
#endif

#endif /* PCL_B200_H_ */
