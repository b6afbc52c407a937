// ft_internal.h -- host-side launch functions shared between the translation units.
#pragma once
#include <cstdlib>
#include <utility>
#include <vector>

#include "ft_device.cuh"


cudaError_t ft_launch_extract_setup(const FtParams& p);
size_t ft_octree_smem_bytes(const FtParams& p, int level);   // dynamic shared memory of k_octree for one level
size_t ft_octree_smem_budget();                              // dynamic shared memory k_octree may use on this device
bool ft_octree_plan(FtLevel& L, size_t smemBudget);          // fills octBinDepth / octCandSmem; false: node list too large
void ft_octree_tables(const FtLevel& L, uint32_t* tabX, uint32_t* tabY);
cudaError_t ft_launch_octree_setup();

// cudaFuncAttributeMaxDynamicSharedMemorySize is per function AND per device: a context must never lower what a live,
// larger context needs. Every kernel with dynamic shared memory gets the device's opt-in maximum (minus its static
// part) once; the launches pass what they actually use.
inline cudaError_t ft_set_max_dynamic_smem(const void* func) {
  int dev = 0, optin = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  if (e != cudaSuccess) return e;
  cudaFuncAttributes fa;
  e = cudaFuncGetAttributes(&fa, func);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - (int)fa.sharedSizeBytes);
}
cudaError_t ft_launch_sbp_setup(const FtParams& p);
cudaError_t ft_launch_stereo_setup(const FtParams& p);
void ft_launch_remap(const FtParams& p, const FtBuffers& b, const uint8_t* rawL, const uint8_t* rawR, const int2* tab, int rawW,
                     int rawH, cudaStream_t st);
void ft_launch_resize_input(const FtParams& p, const FtBuffers& b, const uint8_t* rawL, const uint8_t* rawR, int rawW,
                            cudaStream_t st);
void ft_launch_resize(const FtParams& p, const FtBuffers& b, int level, cudaStream_t st);
void ft_launch_blur(const FtParams& p, const FtBuffers& b, int l0, int l1, cudaStream_t st);
void ft_launch_fast(const FtParams& p, const FtBuffers& b, int l0, int l1, cudaStream_t st);
void ft_launch_octree(const FtParams& p, const FtBuffers& b, int l0, int l1, cudaStream_t st);
void ft_launch_orient_desc(const FtParams& p, const FtBuffers& b, cudaStream_t st);
void ft_launch_stereo_match(const FtParams& p, const FtBuffers& b, const FtStereoBuffers& s, float mbf, float mb,
                            cudaStream_t st);
void ft_launch_fisheye(const FtParams& p, const FtBuffers& b, const FtStereoBuffers& s, const FtCamera& c1,
                       const FtCamera& c2, const FtPose& pose, cudaStream_t st);
void ft_launch_rgbd_depth(const FtParams& p, const FtBuffers& b, const FtStereoBuffers& st, const float2* kpUn, const float* depth,
                          int pitchFloats, float mbf, cudaStream_t s);
void ft_launch_store_scatter(int n, const uint8_t* staged, float* pos, float* normal, float* minmax, uint8_t* desc,
                             cudaStream_t st);
void ft_launch_grid(const FtParams& p, const FtBuffers& b, const FtGridBuffers& g, int fisheye, float minX, float minY,
                    float gridWInv, float gridHInv, const FtUndistort& und, cudaStream_t st);
void ft_launch_gather(const FtParams& p, const FtBuffers& b, const FtGridBuffers& g, const FtStereoBuffers& stb,
                      const FtSbpBuffers& s, const FtFrustumArgs& fa, const FtGatherArgs& ga, int M, cudaStream_t st);
void ft_launch_resolve(const FtBuffers& b, const FtSbpBuffers& s, const FtStereoBuffers& stb, const FtResolveArgs& ra,
                       cudaStream_t st);

// gather -> resolve as one CUDA graph whose kernel-node parameters are refreshed per call (ft_sbp.cu)
struct FtSearchGraph {
  cudaGraphExec_t exec = nullptr;
  cudaGraph_t graph = nullptr;
  cudaGraphNode_t gather = nullptr, resolve = nullptr, seq = nullptr;
};
cudaError_t ft_search_graph_run(FtSearchGraph* G, const FtParams& p, const FtBuffers& b, const FtGridBuffers& g,
                                const FtStereoBuffers& stb, const FtSbpBuffers& s, const FtFrustumArgs& fa,
                                const FtGatherArgs& ga, const FtResolveArgs& ra, int M, cudaStream_t st);
void ft_search_graph_destroy(FtSearchGraph* G);

// Launch `k` as a programmatic dependent of the previous kernel in the stream (FT_PDL=0 in the environment: plain launch).
inline bool ft_pdl_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("FT_PDL"); v = !(e && e[0] == '0'); }
  return v != 0;
}
template <typename... KArgs, typename... Args>
inline cudaError_t ft_launch_pdl(void (*k)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = ft_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, k, std::forward<Args>(args)...);
}

// ---- bag of words (ft_bow.cu) ----
struct ft_vocabulary;
void ft_internal_set_err(const char* msg);   // the thread-local text behind ft_last_error() (ft_context.cu)
int ft_vocabulary_device(const ft_vocabulary* v);
cudaError_t ft_bow_frame_alloc(FtBowFrame* F, int cap, std::vector<void*>& owner);
cudaError_t ft_bow_search_alloc(FtBowSearch* Q, int capF, int capKF, std::vector<void*>& owner);
size_t ft_bow_search_bind(FtBowSearch* Q, int nKF);
int ft_launch_bow_transform(const ft_vocabulary* voc, const FtBowSource& S, const FtBowFrame& F, int maxN, int levelsup,
                            cudaStream_t st);
int ft_launch_bow_search(const FtBowSource& S, const FtBowFrame& F, const FtBowSearch& Q, int nKF, int capF, float nnratio,
                         int checkOri, cudaStream_t st);
