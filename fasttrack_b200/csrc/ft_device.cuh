// ft_device.cuh -- device-side data layout shared by all kernels of the front-end.
//
// Everything a frame needs lives in HBM for the lifetime of the context ("FrameDevice"):
// both eyes' pyramids and blurred pyramids (one slab each, 64-byte row pitch, 256-byte
// level alignment), FAST per-cell candidate slabs, per-level octree outputs, the final
// keypoint/descriptor arrays, stereo results, the 64x48 frame grid and the map-point
// snapshot + candidate lists of the projection search. Kernels receive the geometry as a
// __grid_constant__ FtParams (by value, so it is baked into CUDA-graph nodes).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/fasttrack_b200.h"

#define FT_EDGE_THRESHOLD 19   // reference include/ORBextractor.h:31
#define FT_HALF_PATCH 15       // :30
#define FT_PATCH 31            // :29
#define FT_MIN_BORDER 16       // EDGE_THRESHOLD-3 (ORBextractor.cc:1120)
// Programmatic dependent launch (sm_90+): a kernel launched with the programmatic-stream-serialization attribute may
// start while its predecessor in the stream is still running; FT_PDL_WAIT() blocks until that predecessor has completed
// and its writes are visible (a no-op for a normal launch). Every such kernel executes it on every path before it
// touches anything the predecessor writes. FT_PDL_TRIGGER() in the predecessor lets the dependent be scheduled early.
#define FT_PDL_WAIT() asm volatile("griddepcontrol.wait;" ::: "memory")
#define FT_PDL_TRIGGER() asm volatile("griddepcontrol.launch_dependents;" ::: "memory")
#define FT_OCT_D 12            // quadrant digits of an octree path key (ft_octree.cu)
#define FT_OCT_EVEN_MASK 0x333333u   // digits of even depth are stored complemented

struct FtLevel {
  int w, h;               // level size (cvRound(width * invScale))
  int pitch;              // bytes per row in the slab
  int offset;             // byte offset of the level in the slab
  int nCols, nRows;       // FAST cell grid (ORBextractor.cc:1131-1134)
  int wCell, hCell;
  int maxBorderX, maxBorderY;
  int cellBase;           // first cell id of this level (per eye)
  int cellCap;            // candidate capacity per cell
  int cellKpBase;         // entry offset of this level's cell slab
  int candBase;           // entry offset of this level's flat candidate list
  int candCap;            // capacity of the flat list (worst case NMS survivors)
  int quota;              // mnFeaturesPerLevel
  int nodeCap;            // octree list capacity
  int lvlKpBase;          // entry offset of this level's kept keypoints
  int lvlKpCap;
  int nIni;               // octree roots
  float hX;
  int xTab, yTab;         // offsets into the resize coefficient tables
  int blurTileBase;       // first blur tile id
  int blurTilesX;
  int octBinDepth;        // k_octree: leading quadrant digits that index the bins of its sort
  int octCandSmem;        // k_octree: candidates the level keeps in shared memory (more -> HBM scratch)
  int octTabX, octTabY;   // offsets into the octree path tables
  int octDenseDepth;      // k_octree dense path: depth of the cells k_fast_cells counts into (0 = path off)
  int octDenseBase;       // entry offset of this level's dense cells
  int pad;
};

struct FtParams {
  int nlevels, width, height, nfeatures;
  int iniTh, minTh;
  int totalCells, totalBlurTiles;
  int maxKp;              // capacity of the final per-eye keypoint arrays
  int camType;
  int nEyes;              // 2 = stereo rig, 1 = monocular / RGB-D (only eye 0 is extracted)
  int lap[2][2];
  int umax[16];
  float scale[FT_MAX_LEVELS], invScale[FT_MAX_LEVELS], sigma2[FT_MAX_LEVELS];
  FtLevel lv[FT_MAX_LEVELS];
};

// Per-eye device buffers.
struct FtEye {
  uint8_t* pyr;            // pyramid slab
  uint8_t* blur;           // blurred pyramid slab
  uint32_t* cellKp;        // per-cell candidates: x:12 | y:12 | score:8 (relative to minBorder)
  int* cellCount;          // [totalCells]
  uint32_t* cand;          // flat per-level candidate lists in canonical order (same packing)
  uint8_t* octScratch;     // octree scratch (20 B per candidate slot) for levels whose candidates exceed shared memory
  int* octCnt;             // dense cells of every level: keypoints per cell (filled by k_fast_cells, consumed + zeroed by k_octree)
  unsigned* octBest;       // dense cells: response << 20 | (0xFFFFF - (cell << 9 | slot)) of the best keypoint
  int* lvlCandCount;       // [nlevels]
  uint32_t* lvlKp;         // kept keypoints per level after the octree (level coords, absolute): x:12|y:12|score:8
  int* lvlKpCount;         // [nlevels]
  ft_keypoint* kps;        // final keypoints [maxKp]
  uint8_t* desc;           // final descriptors [maxKp][32]
  int* counts;             // [0]=n, [1]=monoIndex
  long long* octClock;     // debug: clock64 stamps of the octree phases (only with -DFT_OCT_CLOCK)
};

struct FtBuffers {
  FtEye eye[2];
  const int2* xTab;        // per level, per dst column: {sx, a0 | a1<<16}
  const int2* yTab;        // per level, per dst row:    {sy0 | sy1<<16, b0 | b1<<16}
  const uint32_t* octTabX; // per level, per keypoint-area column: root << 24 | x half of the quadtree path (even bits)
  const uint32_t* octTabY; // per level, per keypoint-area row: y half of the quadtree path (odd bits)
  int* status;             // device-side status word (capacity overflow flags)
  unsigned long long* stereoStats;   // FtStereoBuffers::stats (cleared by k_orient_desc for the stereo kernel)
};

// status bits
#define FT_ST_CELL_OVERFLOW 1
#define FT_ST_NODE_OVERFLOW 2
#define FT_ST_KP_OVERFLOW 4
#define FT_ST_CAND_OVERFLOW 8
#define FT_ST_SBP_POOL_OVERFLOW 16
#define FT_ST_RESOLVE_NOCONV 32

__host__ __device__ __forceinline__ uint32_t ft_pack_xys(int x, int y, int s) {
  return (uint32_t)x | ((uint32_t)y << 12) | ((uint32_t)s << 24);
}
__host__ __device__ __forceinline__ int ft_px(uint32_t p) { return (int)(p & 0xFFFu); }
__host__ __device__ __forceinline__ int ft_py(uint32_t p) { return (int)((p >> 12) & 0xFFFu); }
__host__ __device__ __forceinline__ int ft_ps(uint32_t p) { return (int)(p >> 24); }

// 256-bit Hamming distance: __popc over the eight 32-bit words of two uint4 pairs
// (reference ORBmatcher::DescriptorDistance, src/ORBmatcher.cc:2256-2273, computes the same sum).
__device__ __forceinline__ int ft_hamming256(const uint4& a0, const uint4& a1, const uint4& b0, const uint4& b1) {
  return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
         __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

// ---- stereo / projection-search buffers ----
struct FtStereoBuffers {
  float* uRight;           // [maxKp]
  float* depth;            // [maxKp]
  int* bestIdxR;           // coarse match (diagnostic)
  int* sad;                // best SAD per accepted match, -1 otherwise
  int* l2r;                // fisheye [maxKp]
  int* r2l;                // fisheye [maxKp]
  float* p3d;              // fisheye [maxKp][3]
  int* code;               // fisheye per-left code
  unsigned long long* stats;  // [8] counters
};

struct FtCamera {
  int type;
  float p[8];
};

struct FtPose {
  float Rcw[9], tcw[3], Rwc[9], Ow[3];
  float Rlr[9], tlr[3], Rrl[9], trl[3];   // fisheye rig extrinsics (Tlr = T_c1_c2 and its inverse)
};

// per-eye stride of the grid CSR's cell starts: 64*48+1 entries padded to a multiple of 16 bytes (bulk copies)
#define FT_GRID_STRIDE (FT_GRID_COLS * FT_GRID_ROWS + 4)
struct FtGridBuffers {
  int* cellStart;          // [2][FT_GRID_STRIDE] (left, right), 64*48+1 entries used in each
  int* cellIdx;            // [2][maxKp]
  float4* rec;             // [2][maxKp] per keypoint {x, y, uRight (pinhole; -1 otherwise), octave as float bits}
  float2* kpUn;            // [maxKp] mvKeysUn coordinates of the left eye (Frame::UndistortKeyPoints)
};

// Pinhole distortion model of Frame::UndistortKeyPoints (cv::undistortPoints with K, mDistCoef, P = K)
struct FtUndistort {
  int on;                  // mDistCoef[0] != 0 (Frame.cc:773)
  double fx, fy, cx, cy, ifx, ify;
  double k[5];             // k1 k2 p1 p2 k3
};

// Map-point snapshot, frustum scratch, candidate lists and claim tables of the projection search.
struct FtSbpBuffers {
  float* pos;              // [M][3]
  float* normal;           // [M][3]
  float* minmax;           // [M][2] raw mfMinDistance, mfMaxDistance
  uint8_t* desc;           // [M][32]
  int* flags;              // [M] bit0 skip, bit1 Observations()>0
  const int* slot;         // [M] row of pos/normal/minmax/desc for map point j (persistent store), nullptr = j itself
  int* trI;                // [M][4] inView, inViewR, level, levelR
  float* trF;              // [M][9] projX, projY, projXR, depth, viewCos, projXR_r, projYR_r, depthR, viewCosR
  int* listOff;            // [M][2]
  int* listLen;            // [M][2]
  uint32_t* pool;          // candidate entries idx:16 | dist:9 | octave:4
  int poolCap;
  int* cursor;             // [0] pool cursor, [1] nmatches, [2] rounds, [3] non-blocking searched map points, [4] active count,
                           // [5..6..7->5,6 + 7] round flags / pool usage of the last search
  int* active;             // [M] map points with a non-empty candidate list (order irrelevant)
  int* sel;                // [M][2] selected keypoint per branch, -1 none
  int* holderInit;         // [2*maxKp] F.mvpMapPoints on entry (indices), uploaded by the caller
  uint8_t* holderObsInit;  // [2*maxKp]
  int* holder;             // [2*maxKp] result
  uint8_t* holderObs;      // [2*maxKp]
};

// ---- projection-search kernel arguments ----
struct FtFrustumArgs {
  FtCamera cam1, cam2;
  FtPose pose;
  float minX, maxX, minY, maxY;
  float mbf, logScale;
  int nlevels, fisheye;
  float viewCosLimit;
};

struct FtGatherArgs {
  float minX, minY, gridWInv, gridHInv;
  float th; int bFactor; int bFar; float thFar;
  int fisheye;
  int mode;        // 0: local map points (Tracking::SearchLocalPoints), 1: last frame's points (TrackWithMotionModel)
  int direction;   // mode 1: +1 bForward, -1 bBackward, 0 neither (ORBmatcher.cc:1793-1794)
};

struct FtResolveArgs {
  int M, nLeft, nSlots, fisheye;
  float nnratio;
  int mode;        // 1: best match only, no ratio test (ORBmatcher.cc:1847-1880)
  int checkOri;    // mode 1: rotation-histogram consistency (ORBmatcher.cc:1884-1900, 2057-2079)
};


// ---- bag of words (ft_bow.cu): vocabulary tree in child order, per-frame BowVector / FeatureVector, SearchByBoW ----
struct FtVocDevice {
  int L, nNodes;             // depth levels; nodes including the root (node 0)
  int scoring, weighting;    // DBoW2 ScoringType / WeightingType
  const int* firstChild;     // [nNodes] offset of the node's children in the child-order arrays
  const int* nChild;         // [nNodes]
  const int* childNode;      // [nNodes-1] node id of each child slot
  const uint4* childDesc;    // [nNodes-1][2] descriptor of each child slot
  const double* weight;      // [nNodes] node weight (words: idf)
  const int* wordId;         // [nNodes]
  const double* wordWeight;  // [nWords]
};

// where the features of a transform / search come from: the context's device-resident frame (counts read on the
// device, second eye appended for fisheye rigs = the vconcat of Frame.cc:1218) or a plain descriptor array (nFixed)
struct FtBowSource {
  const uint8_t* desc0; const uint8_t* desc1;
  const int* cnt0; const int* cnt1;
  const ft_keypoint* kps0; const ft_keypoint* kps1;
  int nFixed;
};

struct FtBowFrame {
  int* word;                 // [cap] word id per feature
  int* node;                 // [cap] FeatureVector node per feature, -1 = stopped word
  uint32_t* bowIds;          // [cap] BowVector ids ascending
  double* bowVals;           // [cap]
  int* bowStart;             // [cap+1] scratch: run heads of the sorted words
  int* fvIdx;                // [cap] features sorted by (node, index)
  int* fvNode;               // [cap] their nodes
  int* fvGStart;             // [cap+1] group heads
  int* fvMeta;               // [0] features in the FeatureVector, [1] groups
  int* meta;                 // [0] n, [1] nLeft, [2] BowVector size
};

struct FtBowSearch {
  int* result;               // [0] nmatches, [1..3] pad; match follows in the same allocation (one D2H)
  int* match;                // [capF] KeyFrame feature matched to each frame keypoint, -1 none
  int* matchBin;             // [capF] rotation-histogram bin of the match
  int* hist;                 // [32]
  int* kfMeta;               // [0] KeyFrame features in its FeatureVector, [1] groups
  uint8_t* kfBlob;           // uploaded KeyFrame side, one H2D: desc | angle | node | hasMp, bound per call
  uint8_t* kfDesc; float* kfAngle; int* kfNode; uint8_t* kfHasMp;
  int* kfIdxSorted; int* kfNodeSorted; int* kfGStart;
  int kfCap;
};
