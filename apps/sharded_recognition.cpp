// sharded_recognition.cpp — ONE scene over several GPUs from plain C++ (north star: "scene keypoints are sharded
// across the 8 B200s ... the model descriptor library replicated ... correspondences gathered with NCCL"): one host
// thread and one context per GPU, the library's own NCCL communicator (b200_comm_*), every rank calling
// b200_register_scene_shot_sharded; rank 0 holds the scene and receives the result.  The loops it splits are the
// per-keypoint ones of the reference's callback: SHOT at SHOT.cpp:360-371, the matching loop at SHOT.cpp:409-423.
// Parameters as SHOT_scenes.cpp:50-55.
//
// usage: sharded_recognition <model.f32> <model_kp.f32> <scene.f32> <scene_kp.f32> <out_prefix> <ranks> [repeats=1]
// writes <out_prefix>.T (instances x 16 float) and <out_prefix>.corr (b200_corr records); prints the per-call time
#include <b200reg.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <thread>
#include <vector>

static bool load(const char *path, std::vector<float> &xyz) {
  FILE *f = fopen(path, "rb");
  if (!f) {
    fprintf(stderr, "cannot open %s\n", path);
    return false;
  }
  float v[3];
  while (fread(v, sizeof(float), 3, f) == 3) xyz.insert(xyz.end(), v, v + 3);
  fclose(f);
  return true;
}

int main(int argc, char **argv) {
  if (argc < 7) {
    fprintf(stderr, "usage: %s model.f32 model_kp.f32 scene.f32 scene_kp.f32 out_prefix ranks [repeats]\n", argv[0]);
    return 2;
  }
  std::vector<float> model, model_kp, scene, scene_kp;
  if (!load(argv[1], model) || !load(argv[2], model_kp) || !load(argv[3], scene) || !load(argv[4], scene_kp)) return 1;
  const std::string prefix = argv[5];
  const int world = atoi(argv[6]);
  const int repeats = argc > 7 ? atoi(argv[7]) : 1;
  if (world < 1 || world > 64 || repeats < 1) return 2;

  b200_shot_params p;
  p.normal_k = 20;
  p.normal_radius = 0.0;
  p.descr_radius = 0.02;
  p.match_mode = 1;
  p.match_thr = 0.25f;
  p.gc_size = 0.02;
  p.gc_threshold = 2;
  p.max_instances = 4096;

  char token[128] = {0};
  if (world > 1 && b200_comm_unique_id(token, sizeof(token)) != B200_OK) {
    fprintf(stderr, "%s\n", b200_last_error(nullptr));
    return 1;
  }
  const int Ks = (int)scene_kp.size() / 3;
  std::vector<float> T((size_t)p.max_instances * 16);
  std::vector<int> off((size_t)p.max_instances + 1);
  std::vector<b200_corr> ic((size_t)(Ks > 0 ? Ks : 1)), co((size_t)(Ks > 0 ? Ks : 1));
  int n_inst = 0, n_corr = 0;
  double ms_per_call = 0.0;
  std::vector<int> status((size_t)world, B200_OK);

  auto rank_main = [&](int r) {
    b200_ctx *ctx = nullptr;
    b200_model *m = nullptr;
    int rc = b200_ctx_create(&ctx, r, nullptr);
    if (rc == B200_OK && world > 1) rc = b200_comm_init(ctx, token, r, world);
    if (rc == B200_OK)
      rc = b200_model_create_shot(ctx, model.data(), (int)model.size() / 3, 3, model_kp.data(), (int)model_kp.size() / 3, 3,
                                  &p, &m);   // every rank keeps its own copy of the library
    for (int it = 0; rc == B200_OK && it < repeats + 1; ++it) {   // the first call warms the arena up
      const auto t0 = std::chrono::steady_clock::now();
      if (r == 0)
        rc = b200_register_scene_shot_sharded(ctx, m, 0, scene.data(), (int)scene.size() / 3, 3, scene_kp.data(), Ks, 3, &p,
                                              T.data(), off.data(), ic.data(), (int)ic.size(), &n_inst, co.data(), &n_corr);
      else
        rc = b200_register_scene_shot_sharded(ctx, m, 0, nullptr, 0, 3, nullptr, 0, 3, &p, nullptr, nullptr, nullptr, 0,
                                              nullptr, nullptr, nullptr);
      if (rc == B200_ERR_CAPACITY) rc = B200_OK;
      if (r == 0 && it > 0)
        ms_per_call += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() / repeats;
    }
    if (rc != B200_OK) fprintf(stderr, "rank %d: %s\n", r, b200_last_error(ctx));
    status[(size_t)r] = rc;
    if (m) b200_model_destroy(m);
    if (ctx) b200_ctx_destroy(ctx);
  };
  std::vector<std::thread> threads;
  for (int r = 0; r < world; ++r) threads.emplace_back(rank_main, r);
  for (auto &t : threads) t.join();
  for (int r = 0; r < world; ++r)
    if (status[(size_t)r] != B200_OK) return 1;

  printf("%d rank(s): %d correspondences, %d instances, %.3f ms per scene\n", world, n_corr, n_inst, ms_per_call);
  const int kept = n_inst < p.max_instances ? n_inst : p.max_instances;
  FILE *f = fopen((prefix + ".T").c_str(), "wb");
  fwrite(T.data(), sizeof(float), (size_t)kept * 16, f);
  fclose(f);
  f = fopen((prefix + ".corr").c_str(), "wb");
  fwrite(co.data(), sizeof(b200_corr), (size_t)n_corr, f);
  fclose(f);
  return 0;
}
