/*
 * hfnet_b200.h -- C-ABI of libhfnet_b200.so: a B200 (sm_100a) implementation of the HFNet-SLAM per-frame front-end.
 *
 * Plain C: opaque handles, plain pointers and sizes, integer status codes.  No torch / OpenCV / Eigen types.
 * Every entry point cites the reference interface (LiuLimingCode/HFNet_SLAM, paths relative to the repository
 * root) it replaces.  The reference-side bindings (C++ shim classes) are in include/HFNetB200Model.h and
 * INTEGRATION.md.
 *
 * Memory spaces: functions without a suffix take HOST pointers and are synchronous (inputs copied in, results
 * copied out, stream synchronised) -- these are the drop-in calls.  Functions ending in `_dev` take DEVICE pointers
 * (same CUDA device as the context), enqueue on the context's stream and return without synchronising; call
 * hfb_sync() before reading results.  Calls on one context must be serialised by the caller; distinct contexts are
 * fully concurrent (own stream, own workspace) -- the reference calls Detect concurrently on different model
 * objects only (src/Extractors/HFextractor.cc:228-243,265).
 *
 * Errors: 0 = HFB_OK; otherwise hfb_last_error(ctx) holds a message.  Nothing throws, nothing calls exit()
 * (the reference returns bool from Detect, src/Extractors/HFNetRTModel.cc:87-107, and exit(-1)s on load failure,
 * src/Extractors/BaseModel.cc:140).
 */
#ifndef HFNET_B200_H_
#define HFNET_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HFB_VERSION 100

enum hfb_status {
  HFB_OK = 0,
  HFB_ERR_INVALID = 1,   /* bad argument / wrong mode (reference: Detect returns false) */
  HFB_ERR_CUDA = 2,      /* CUDA runtime failure */
  HFB_ERR_STATE = 3,     /* e.g. weights not loaded */
  HFB_ERR_CAPACITY = 4,  /* a caller- or context-side capacity was exceeded */
  HFB_ERR_NUMERIC = 5    /* e.g. reduced camera system not positive definite */
};

#define HFB_DESC_DIM 256     /* local descriptor width, src/Extractors/HFNetRTModel.cc:181 */
#define HFB_GLOBAL_DIM 4096  /* global descriptor width, src/Extractors/HFNetRTModel.cc:201 */
#define HFB_MAX_LEVELS 8

typedef struct hfb_ctx hfb_ctx;

/* ------------------------------------------------------------------------------------------------------ context
 * Replaces InitAllModels(strModelPath, type, ImSize, nLevels, scaleFactor) (src/Extractors/BaseModel.cc:24-93):
 * one context holds the plans of ALL pyramid levels (level l is cvRound(H*s) x cvRound(W*s), s = 1/scaleFactor^l)
 * plus the matcher / database / BA workspaces. */
typedef struct hfb_config {
  int32_t device;         /* CUDA ordinal */
  int32_t height, width;  /* level-0 image size (any size; the network crops to multiples of 8, hf_net.py:188-190) */
  int32_t n_levels;       /* 1 = bare Detect (ExtractSingleLayer, HFextractor.cc:175-182) */
  float scale_factor;     /* Extractor.scaleFactor, src/Settings.cc:443-470 */
  int32_t max_keypoints;  /* capacity per level and frame */
  int32_t max_batch;      /* frames per hfb_extract_batch call (>= 1) */
  int32_t with_global;    /* 1: level 0 also produces the 4096-d global descriptor (kImageToLocalAndGlobal) */
} hfb_config;

int hfb_create(const hfb_config* cfg, hfb_ctx** out);
void hfb_destroy(hfb_ctx* ctx);
const char* hfb_last_error(const hfb_ctx* ctx);
int hfb_version(void);
int hfb_sync(hfb_ctx* ctx);
/* Raw cudaStream_t of the context (for callers that enqueue their own work / time with events). */
void* hfb_stream(hfb_ctx* ctx);
/* Number of kernels this library has launched on ctx since creation (graph replays count their kernel nodes). */
uint64_t hfb_launch_count(const hfb_ctx* ctx);

/* Page-locked host memory for zero-copy staging: images / output arrays that live in memory obtained here are DMA'd
 * directly (no intermediate copy).  Ordinary (pageable) pointers are accepted everywhere and are staged internally. */
void* hfb_host_alloc(size_t bytes);
void hfb_host_free(void* p);

/* Weights: flat 'HFB2WTS1' blob (hfnet_slam_b200/weights.py), BatchNorm folded.  Replaces the ONNX parse + TensorRT
 * engine build of HFNetRTModel::LoadHFNetTRModel (src/Extractors/HFNetRTModel.cc:208-254). */
int hfb_load_weights(hfb_ctx* ctx, const void* blob, size_t nbytes);

/* ------------------------------------------------------------------------------------------------------ extraction
 * Caller-allocated SoA outputs with capacity sum(n_per_level).  KeyPoint{pt.x, pt.y, response, octave}
 * (src/Extractors/HFNetRTModel.cc:150-166, HFextractor.cc:272-279); descriptors row-major N x 256, unit L2 rows,
 * contiguous (src/Matcher.cc:843-844 requires isContinuous()); global descriptor 4096 floats. */
typedef struct hfb_features {
  float* x;                 /* [cap] column * scaleFactor^level */
  float* y;                 /* [cap] row    * scaleFactor^level */
  float* response;          /* [cap] */
  int32_t* octave;          /* [cap] */
  float* descriptors;       /* [cap * 256], or NULL: the local descriptors stay resident in HBM (the association, the
                               windowed searches on resident frames and the keyframe store read them there; fetch later
                               with hfb_fetch_features if a caller needs them on the host) */
  float* global_descriptor; /* [4096] or NULL */
  int32_t n_per_level[HFB_MAX_LEVELS]; /* out: keypoints found per level (<= requested) */
  int32_t n_total;          /* out */
} hfb_features;

/* HFextractor::operator() (src/Extractors/HFextractor.cc:142-157) == pyramid (:159-173) + BaseModel::Detect per level
 * (src/Extractors/HFNetRTModel.cc:84-110) + concat (:272-281).  image: CV_8UC1, `stride` bytes per row.
 * n_per_level[l]: keypoint budget of level l (HFextractor.cc:108-119).  threshold: score >= threshold. */
int hfb_extract(hfb_ctx* ctx, const uint8_t* image, int32_t height, int32_t width, int32_t stride,
                const int32_t* n_per_level, float threshold, hfb_features* out);

/* BaseModel::Detect of ONE pyramid level on a context that holds all levels (src/Extractors/HFNetRTModel.cc:84-110): the
 * unmodified HFextractor hands every per-level model object its own pre-scaled image (HFextractor.cc:228-243).  `image`
 * is level `level`'s image (cvRound(H / s^level) x cvRound(W / s^level)); keypoints come back in LEVEL coordinates with
 * octave 0, the caller applies octave and scale (HFextractor.cc:272-279).  The global descriptor is produced by level 0
 * only (kImageToLocalAndGlobal) and only when out->global_descriptor is non-NULL. */
int hfb_extract_level(hfb_ctx* ctx, int32_t level, const uint8_t* image, int32_t height, int32_t width, int32_t stride,
                      int32_t n_keypoints, float threshold, hfb_features* out);

/* Same for `n_images` independent frames (one per camera stream); images[i] are host pointers. */
int hfb_extract_batch(hfb_ctx* ctx, const uint8_t* const* images, int32_t n_images, int32_t stride,
                      const int32_t* n_per_level, float threshold, hfb_features* outs);

/* Device-resident variant: d_images = n_images contiguous [height*width] u8 frames already in HBM.  Results stay in
 * context-owned device buffers; fetch with hfb_fetch_features (synchronises). */
int hfb_extract_batch_dev(hfb_ctx* ctx, const uint8_t* d_images, int32_t n_images, const int32_t* n_per_level,
                          float threshold);
int hfb_fetch_features(hfb_ctx* ctx, int32_t image_index, hfb_features* out);

/* Per-launch device timing of one un-graphed extraction of the frames already resident in the context (after
 * hfb_extract_batch*): JSON array of {"name","ms","bytes","flops"}; bytes / flops are the algorithmic figures of
 * DESIGN.md.  Replaces the reference's REGISTER_TIMES stage timers (src/Frame.cc:311-319). */
int hfb_profile_extract(hfb_ctx* ctx, int32_t n_images, const int32_t* n_per_level, float threshold, char* json_out,
                        size_t cap);

/* Network-tail stages on caller-supplied dense maps (parity hooks; also the reference's CPU stages):
 * hfb_nms            == simple_nms(radius 4, iterations 2) (hfnet/models/utils/layers.py:10-32)
 * hfb_select_sample  == GetLocalFeaturesFromTensor (src/Extractors/HFNetRTModel.cc:139-196): threshold scan,
 *                       top-k by response (ties: column-major scan order), warp, Resampler
 *                       (src/Extractors/BaseModel.cc:491-562), cv::normalize per row.
 * hfb_resize_linear_u8 == cv::resize(..., INTER_LINEAR) on CV_8UC1 (HFextractor.cc:170), bit-exact. */
int hfb_nms(hfb_ctx* ctx, const float* scores, int32_t height, int32_t width, float* scores_nms);
int hfb_select_sample(hfb_ctx* ctx, const float* scores_nms, int32_t height, int32_t width, const float* desc_map,
                      int32_t desc_h, int32_t desc_w, int32_t n_keypoints, float threshold, float* x, float* y,
                      float* response, float* descriptors, int32_t* n_out);
int hfb_resize_linear_u8(hfb_ctx* ctx, const uint8_t* src, int32_t sh, int32_t sw, uint8_t* dst, int32_t dh,
                         int32_t dw);

/* Frame's calibration (mK, mDistCoef; src/Frame.cc:760-825).  K = (fx, fy, cx, cy); dist = OpenCV's (k1, k2, p1, p2[, k3
 * [, k4, k5, k6[, s1, s2, s3, s4]]]), n_dist in {0, 4, 5, 8, 12}.  With dist[0] != 0 every extraction also leaves the
 * undistorted keypoint coordinates (Frame::mvKeysUn) resident next to the distorted ones, and the searches on resident
 * frames (hfb_match_projection_frame) use them -- as the reference's do.  dist[0] == 0 (or n_dist == 0) is the
 * reference's early return: mvKeysUn = mvKeys (:762-766).
 * hfb_undistort_points   == cv::undistortPoints(mat, mat, K, mDistCoef, cv::Mat(), mK) on N x 2 floats (:778), bit-exact
 *                           with OpenCV 4's 5-iteration fixed point in double.
 * hfb_fetch_undistorted  : mvKeysUn coordinates of frame `image_index` of the last extraction (first n keypoints).
 * hfb_image_bounds       == Frame::ComputeImageBounds (:796-825): {mnMinX, mnMaxX, mnMinY, mnMaxY}. */
int hfb_set_camera(hfb_ctx* ctx, const float* K, const float* dist, int32_t n_dist);
int hfb_undistort_points(hfb_ctx* ctx, const float* x, const float* y, int32_t n, float* x_un, float* y_un);
int hfb_fetch_undistorted(hfb_ctx* ctx, int32_t image_index, float* x_un, float* y_un, int32_t n);
int hfb_image_bounds(hfb_ctx* ctx, int32_t width, int32_t height, float* bounds);

/* Debug / parity hook: the tensor-core GEMM primitive on caller data (fp16 operands, fp32 result [B*H*W][N]).
 * A: [B*H*W][K] (NHWC when conv3x3), Wt: [N][conv3x3 ? 9K : K]; use_tc = 0 runs the CUDA-core cross-check kernel. */
int hfb_debug_gemm(hfb_ctx* ctx, const float* A, int B, int H, int W, int K, const float* Wt, int N, const float* bias,
                   int relu6, int conv3x3, int use_tc, int BN, float* out);

/* Debug / parity hook: copy an intermediate tensor of the LAST extraction to host.  name is one of
 * "layer_1".."layer_18", "desc_conv1", "det_conv1", "det_logits", "scores_dense", "scores_dense_nms",
 * "local_descriptor_map", "vlad_norm", "global_descriptor", "pyramid".  Returns element count in *n. */
int hfb_debug_tensor(hfb_ctx* ctx, const char* name, int32_t image_index, int32_t level, float* out, size_t cap,
                     size_t* n, int32_t* dims4);

/* ------------------------------------------------------------------------------------------------------ matching
 * Dense mutual-nearest-neighbour matching of 256-d descriptors (A: na x 256, B: nb x 256, row-major fp32).
 * Outputs are dense per-row-of-A arrays: match_idx[i] = j or -1, match_val[i] = distance or cosine.
 *
 * hfb_match_mutual_l2  == cv::BFMatcher(NORM_L2, crossCheck=true).match + `dist < max_dist`
 *                         (Matcher::SearchByBoW, src/Matcher.cc:220-263, :561-621; TH_LOW = 0.6)
 * hfb_match_mutual_cos == sgemm D1*D2^T + row argmax above `min_cos` + column cross-check
 *                         (Matcher::SearchForTriangulation, src/Matcher.cc:845-889; floor 1-0.5*TH_HIGH^2 = 0.71875,
 *                         strict '>', lowest index wins ties) */
int hfb_match_mutual_l2(hfb_ctx* ctx, const float* A, int32_t na, const float* B, int32_t nb, float max_dist,
                        int32_t* match_idx, float* match_val, int32_t* n_matches);
int hfb_match_mutual_cos(hfb_ctx* ctx, const float* A, int32_t na, const float* B, int32_t nb, float min_cos,
                         int32_t* match_idx, float* match_val, int32_t* n_matches);

/* Batched: pair p matches rows [a_off[p], a_off[p]+a_cnt[p]) of A_all against rows [b_off[p], b_off[p]+b_cnt[p]) of
 * B_all (one CreateNewMapPoints = <= 30 neighbour keyframes, src/LocalMapping.cc:516-519).  match_idx/match_val are
 * laid out like A's rows of each pair, concatenated in pair order (total = sum a_cnt); indices are pair-local.
 * mode: 0 = l2 (thr = max_dist), 1 = cos (thr = min_cos). */
int hfb_match_batch(hfb_ctx* ctx, int32_t mode, const float* A_all, int32_t na_total, const float* B_all,
                    int32_t nb_total, int32_t n_pairs, const int32_t* a_off, const int32_t* a_cnt,
                    const int32_t* b_off, const int32_t* b_cnt, float thr, int32_t* match_idx, float* match_val);
/* Device-pointer variant: descriptors, pair table and outputs all in HBM; enqueues on the context stream, no sync.
 * d_pair_tab = int32[4][n_pairs] holding a_off | a_cnt | b_off | b_cnt; max_*_cnt bound the per-pair counts (they
 * size the launch grid).  d_match_idx / d_match_val must hold na_total entries. */
int hfb_match_batch_dev(hfb_ctx* ctx, int32_t mode, const float* dA_all, int32_t na_total, const float* dB_all,
                        int32_t nb_total, int32_t n_pairs, const int32_t* d_pair_tab, int32_t max_a_cnt,
                        int32_t max_b_cnt, float thr, int32_t* d_match_idx, float* d_match_val);

/* Windowed matching: the descriptor stage of Matcher::SearchByProjection (src/Matcher.cc:40-210, 1574-1721) +
 * Frame::GetFeaturesInArea (src/Frame.cc:659-725).  For every query (projected map point: descriptor, (u, v), window
 * radius r, octave range [min_level, max_level], max_level < 0 = unbounded) the HFB_PROJ_TOPK = 4 nearest frame
 * features with |x-u| < r, |y-v| < r, octave in range and f_skip[j] == 0, sorted by L2 distance (ties: lower index).
 * Outputs are [nq][4]: feature index or -1, distance or FLT_MAX, octave or -1.  The caller replays the reference's
 * sequential bookkeeping (skip features claimed by earlier map points, best / second + level-aware ratio test) on
 * these lists; see hfnet_slam_b200/matcher.py. */
#define HFB_PROJ_TOPK 4
int hfb_match_projection(hfb_ctx* ctx, const float* Q, int32_t nq, const float* q_uv, const float* q_radius,
                         const int32_t* q_min_level, const int32_t* q_max_level, const float* F, int32_t nf,
                         const float* f_xy, const int32_t* f_level, const uint8_t* f_skip, int32_t* cand_idx,
                         float* cand_dist, int32_t* cand_level);
/* Same with Matcher::Fuse's reprojection gate (src/Matcher.cc:1160-1190): a candidate is skipped iff
 * ((u - x)^2 + (v - y)^2) * f_inv_sigma2[candidate] > chi2_max (f_inv_sigma2 = mvInvLevelSigma2 of the feature's octave,
 * chi2_max = 5.99 mono).  f_inv_sigma2 == NULL: no gate (== hfb_match_projection). */
int hfb_match_projection_gated(hfb_ctx* ctx, const float* Q, int32_t nq, const float* q_uv, const float* q_radius,
                               const int32_t* q_min_level, const int32_t* q_max_level, const float* F, int32_t nf,
                               const float* f_xy, const int32_t* f_level, const uint8_t* f_skip,
                               const float* f_inv_sigma2, float chi2_max, int32_t* cand_idx, float* cand_dist,
                               int32_t* cand_level);

/* The same search on RESIDENT descriptors (SearchByProjection(CurrentFrame, LastFrame), src/Matcher.cc:
 * 1574-1650, without moving descriptors): features = frame `frame_index` of the last hfb_extract* call as it sits in HBM
 * (its first nf keypoints), queries = rows q_prev_index[i] of the previous frame of that stream (frame_index - 1 of the
 * same call, or the frame carried over from the previous call).  Only the query windows are uploaded. */
int hfb_match_projection_frame(hfb_ctx* ctx, int32_t frame_index, const int32_t* q_prev_index, int32_t nq,
                               const float* q_uv, const float* q_radius, const int32_t* q_min_level,
                               const int32_t* q_max_level, int32_t nf, const uint8_t* f_skip, const float* f_inv_sigma2,
                               float chi2_max, int32_t* cand_idx, float* cand_dist, int32_t* cand_level);

/* Tracking's frame-to-previous-frame descriptor association (the brute-force stage behind
 * Matcher::SearchByBoW / SearchForInitialization call sites, src/Tracking.cc:2030,1796, which match the current frame
 * against mLastFrame / the reference keyframe) with the descriptors of the last hfb_extract_batch* still resident in
 * HBM.  The association is STREAMING: every frame is matched against the previous frame of its stream, also across
 * calls -- the context keeps the descriptors it needs from the previous extraction.
 *   stream mode 0 (default): a batch holds B consecutive frames of ONE stream; frame b is matched against frame b-1,
 *                            frame 0 against the last frame of the previous call (a batch of 1 == the reference's loop).
 *   stream mode 1:           a batch holds one frame of each of B independent streams (cameras); frame b is matched
 *                            against frame b of the previous call (same batch size).
 * A frame without a predecessor (first call, after hfb_reset_stream, after a mode change) has no matches.  No sync. */
int hfb_match_consecutive_dev(hfb_ctx* ctx, int32_t n_images, int32_t mode, float thr);
int hfb_set_stream_mode(hfb_ctx* ctx, int32_t mode);
/* Forget the previous frame(s): Tracking::Reset / ResetActiveMap start again from an empty mLastFrame
 * (src/Tracking.cc:3256-3330). */
int hfb_reset_stream(hfb_ctx* ctx);
/* Extraction and the association above as ONE call: the matching kernels only need the local features, so they (and,
 * for page-locked batch-contiguous outputs, the transfer of keypoints / descriptors / match rows) are enqueued on the
 * main stream while the global branch (layer_8 .. FC) is still computing on the side stream.  match_mode < 0 skips the
 * association (== hfb_extract_batch).  match_idx / match_val: [n_images][kp_cap] as for hfb_match_consecutive. */
int hfb_extract_match_batch(hfb_ctx* ctx, const uint8_t* const* images, int32_t n_images, int32_t stride,
                            const int32_t* n_per_level, float threshold, hfb_features* outs, int32_t match_mode,
                            float match_thr, int32_t* match_idx, float* match_val);
int hfb_extract_match_batch_dev(hfb_ctx* ctx, const uint8_t* d_images, int32_t n_images, const int32_t* n_per_level,
                                float threshold, int32_t match_mode, float match_thr);
/* Same, synchronous, results to host: match_idx / match_val are [n_images][kp_cap] rows (kp_cap = n_levels *
 * max_keypoints; row b holds frame b's matches into the keypoints of its stream's previous frame, -1 = unmatched).  This is the call
 * Tracking makes right after Frame construction: the previous frame's descriptors are still resident, so nothing is
 * uploaded (Matcher::SearchByBoW with host cv::Mat descriptors re-reads both sets, src/Matcher.cc:220-263). */
int hfb_match_consecutive(hfb_ctx* ctx, int32_t n_images, int32_t mode, float thr, int32_t* match_idx,
                          float* match_val);
/* match_idx / match_val of frame `image_index` (indices into the previous frame's keypoints), first n rows. */
int hfb_fetch_matches(hfb_ctx* ctx, int32_t image_index, int32_t* match_idx, float* match_val, int32_t n);

/* Keyframe descriptor store: the local descriptors of keyframes kept RESIDENT in HBM (fp32 rows + the split-precision image
 * of the tensor-core contraction, prepared once), so that LocalMapping's per-keyframe matching -- the current keyframe
 * against <= 30 covisible keyframes in CreateNewMapPoints / SearchInNeighbors (src/LocalMapping.cc:513-893,
 * Matcher::SearchForTriangulation src/Matcher.cc:763-936) -- moves no descriptors: keyframe ids go up, match rows come
 * back.  A store lives on one device and may be filled from one context (Tracking's, straight from the extraction that
 * produced the keyframe: hfb_kfstore_put_frame) and read from another (LocalMapping's).  KeyFrame::SetBadFlag -> erase. */
typedef struct hfb_kfstore hfb_kfstore;
int hfb_kfstore_create(hfb_ctx* ctx, int32_t n_slots, int32_t rows_per_slot, hfb_kfstore** out);
void hfb_kfstore_destroy(hfb_kfstore* store);
int hfb_kfstore_put_frame(hfb_ctx* ctx, hfb_kfstore* store, int64_t kf_id, int32_t frame_index, int32_t n);
int hfb_kfstore_put(hfb_ctx* ctx, hfb_kfstore* store, int64_t kf_id, const float* descriptors, int32_t n);
int hfb_kfstore_erase(hfb_kfstore* store, int64_t kf_id);
int32_t hfb_kfstore_size(hfb_kfstore* store);
/* Keyframe kf_a against the n_b stored keyframes kf_b[] as one batched launch (mode / thr as hfb_match_batch).
 * match_idx / match_val: [n_b][rows of kf_a] (*rows_a_out), indices into the rows of the respective neighbour. */
int hfb_match_kf_neighbours(hfb_ctx* ctx, hfb_kfstore* store, int64_t kf_a, const int64_t* kf_b, int32_t n_b, int32_t mode,
                            float thr, int32_t* match_idx, float* match_val, int32_t* rows_a_out);

/* MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:331-400) for a ragged batch of map points: point p owns the
 * descriptor rows offsets[p] .. offsets[p+1]-1 (its observations, <= 128); best_index[p] is the row (relative to the
 * point's first) whose median L2 distance to the point's other descriptors is smallest (first such row, the
 * reference's strict '<' scan), best_median[p] that median (sorted[int(0.5 * (N - 1))]); -1 for a point without rows. */
int hfb_distinctive_descriptors(hfb_ctx* ctx, const float* descriptors, const int32_t* offsets, int32_t n_points,
                                int32_t* best_index, float* best_median);

/* ------------------------------------------------------------------------------------------------------ keyframe DB
 * Replaces KeyFrameDatabase's linear scan (src/KeyFrameDatabase.cc:75-256).  Rows live in HBM as fp32 [capacity][dim].
 * The covisibility accumulation stays with the caller (it walks the KeyFrame graph); see include/HFNetB200Model.h. */
typedef struct hfb_kfdb hfb_kfdb;

int hfb_kfdb_create(hfb_ctx* ctx, int32_t dim, int32_t capacity, hfb_kfdb** out);
void hfb_kfdb_destroy(hfb_kfdb* db);
/* KeyFrameDatabase::add / erase / clear (src/KeyFrameDatabase.cc:31-52). ids are the caller's KeyFrame::mnId. */
int hfb_kfdb_add(hfb_kfdb* db, const int64_t* ids, const float* descriptors, int32_t n);
int hfb_kfdb_add_dev(hfb_kfdb* db, const int64_t* ids /*host*/, const float* d_descriptors, int32_t n);
int hfb_kfdb_erase(hfb_kfdb* db, int64_t id);
int hfb_kfdb_clear(hfb_kfdb* db);
/* KeyFrameDatabase::clearMap(Map*) (src/KeyFrameDatabase.cc:54-68): add with the keyframe's map recorded (map_ids[i] =
 * any stable identifier of KeyFrame::GetMap(); hfb_kfdb_add records map 0), then drop every keyframe of one map. */
int hfb_kfdb_add_tagged(hfb_kfdb* db, const int64_t* ids, const int64_t* map_ids, const float* descriptors, int32_t n);
int hfb_kfdb_clear_map(hfb_kfdb* db, int64_t map_id);
int32_t hfb_kfdb_size(const hfb_kfdb* db);
/* One query (src/KeyFrameDatabase.cc:85-104 / :177-192): score_i = max(0, 1 - ||q - d_i||_2) for every row,
 * best = max score, candidates = { i : score_i > max(floor, rel * best) } (strict).  rel = 0.8; floor = 0 for
 * DetectNBestCandidates, 0.5 for DetectRelocalizationCandidates.  Candidates are returned sorted by ascending id.
 * *n_cand may exceed cap (then only cap are written and HFB_ERR_CAPACITY is returned). */
int hfb_kfdb_query(hfb_kfdb* db, const float* query, float rel, float floor, int64_t* cand_ids, float* cand_scores,
                   int32_t cap, int32_t* n_cand, float* best_score);
/* n_queries queries as ONE pass over the rows (relocalisation of several frames / sessions at once; BASELINE.json
 * configs[3] with Q = 64): the contraction q . d runs on the tensor cores (kind::tf32) to SELECT the pairs that can be
 * candidates or the best row under its error bound, those pairs are re-scored exactly, and candidate sets / best scores
 * are decided on the exact values -- the results equal n_queries calls of hfb_kfdb_query.  Outputs are [n_queries][cap]
 * (ids ascending per query), n_cand[q] may exceed cap (then HFB_ERR_CAPACITY is returned after filling what fits). */
int hfb_kfdb_query_batch(hfb_kfdb* db, const float* queries, int32_t n_queries, float rel, float floor, int32_t cap,
                         int64_t* cand_ids, float* cand_scores, int32_t* n_cand, float* best_scores);
/* Device-resident queries, enqueue only (throughput measurement): results stay in the database's device buffers. */
int hfb_kfdb_query_batch_dev(hfb_kfdb* db, const float* d_queries, int32_t n_queries, float rel, float floor);
/* Scores of arbitrary keyframes under the LAST query (mPlaceRecognitionScore of covisible neighbours,
 * src/KeyFrameDatabase.cc:117-131).  Unknown ids get -1. */
int hfb_kfdb_scores_of(hfb_kfdb* db, const int64_t* ids, int32_t n, float* scores);
/* Scan only, device-resident: d_query [n_queries][dim] -> d_scores [n_queries][size] and d_best [n_queries]
 * (max score as float).  No sync.  Used for the HBM-resident throughput measurement and by the sharded database. */
int hfb_kfdb_scan_dev(hfb_kfdb* db, const float* d_query, int32_t n_queries, float* d_scores, float* d_best);
/* Sharded database (row-shard by id % world): fixed-size record of this shard for ONE all-gather:
 * record = { float local_best; int32 count; int32 overflow; int32 pad; { float score; int32 pad; int64 id }[k] }
 * holding the top-k (by score, ties by id) of { i : score_i > max(floor, rel * local_best) } -- a superset of the
 * shard's share of the global candidate set because local_best <= global_best.  Bytes = 16 + 16*k. */
int hfb_kfdb_query_shard(hfb_kfdb* db, const float* query, float rel, float floor, int32_t k, void* record);
/* The same exchange without the host in the loop: every rank's scan is followed by ONE single-CTA kernel that builds the
 * record on the device, stores it into every peer's inbox over NVLink (peer memory mapped with CUDA IPC), publishes a
 * flag, waits for the peers' flags and merges by the global best -- the host sees one enqueue and one small D2H.
 *   hfb_kfdb_shard_setup          allocates this rank's inbox for `world` records of k entries; ipc_handle_out receives the
 *                                 64-byte cudaIpcMemHandle_t the other ranks need (NULL for same-process use)
 *   hfb_kfdb_shard_connect        multi-process: the handles of ranks 0 .. world-1, [world][64] bytes, gathered by the
 *                                 caller (e.g. one torch.distributed all_gather at start-up)
 *   hfb_kfdb_shard_connect_local  same process: the shard objects of ranks 0 .. world-1
 *   hfb_kfdb_query_sharded        collective: all ranks call it with the same query, in the same order; returns the GLOBAL
 *                                 candidate set (ids ascending) and best score on every rank; *overflow = a shard had more
 *                                 than k rows above the global bar.  _begin enqueues, _end synchronises and copies out. */
int hfb_kfdb_shard_setup(hfb_kfdb* db, int32_t rank, int32_t world, int32_t k, void* ipc_handle_out);
int hfb_kfdb_shard_connect(hfb_kfdb* db, const void* all_handles);
int hfb_kfdb_shard_connect_local(hfb_kfdb* db, hfb_kfdb* const* shards);
int hfb_kfdb_query_sharded_begin(hfb_kfdb* db, const float* query, float rel, float floor);
int hfb_kfdb_query_sharded_end(hfb_kfdb* db, int64_t* cand_ids, float* cand_scores, int32_t cap, int32_t* n_cand,
                               float* best_score, int32_t* overflow);
int hfb_kfdb_query_sharded(hfb_kfdb* db, const float* query, float rel, float floor, int64_t* cand_ids, float* cand_scores,
                           int32_t cap, int32_t* n_cand, float* best_score, int32_t* overflow);

/* ------------------------------------------------------------------------------------------------------ local BA
 * Replaces the numeric core of Optimizer::LocalBundleAdjustment (src/Optimizer.cc:1116-1498): everything between
 * optimizer.initializeOptimization() and the outlier test, i.e. g2o's optimize(10) with BlockSolver_6_3 +
 * Levenberg + Huber(sqrt 5.991) on EdgeSE3ProjectXYZ edges.  Per-edge residual/Jacobian (include/OptimizableTypes.h:
 * 99-110, src/OptimizableTypes.cpp:139-159), Hessian assembly (Thirdparty/g2o/g2o/core/base_binary_edge.hpp:55-121)
 * and the Schur complement + landmark back-substitution (Thirdparty/g2o/g2o/core/block_solver.hpp:354-486) run on the
 * GPU in fp64; only the reduced 6n x 6n camera system is factorised on the host.  Edges MUST be sorted by point. */
typedef struct hfb_lba_problem {
  int32_t n_cams, n_points, n_edges;
  const double* poses;       /* [n_cams*7] Tcw as qx qy qz qw tx ty tz (g2o::SE3Quat) */
  const uint8_t* fixed;      /* [n_cams] 1 = fixed keyframe (src/Optimizer.cc:1240-1262) */
  const double* points;      /* [n_points*3] world positions */
  const int32_t* edge_cam;   /* [n_edges] */
  const int32_t* edge_point; /* [n_edges] non-decreasing */
  const double* obs;         /* [n_edges*2] undistorted keypoint (kpUn.pt) */
  const double* inv_sigma2;  /* [n_edges] mvInvLevelSigma2[octave] (src/Optimizer.cc:1316-1317) */
  float K[4];                /* fx fy cx cy (Pinhole::mvParameters are float, src/CameraModels/Pinhole.cpp:35-41) */
  double huber_delta;        /* thHuberMono = sqrt(5.991), src/Optimizer.cc:1206 */
} hfb_lba_problem;

typedef struct hfb_lba_stats {
  int32_t iterations;   /* outer LM iterations executed */
  int32_t trials;       /* linear solves (LM trials) */
  double initial_chi2;  /* robust chi2 before */
  double final_chi2;    /* robust chi2 after */
  double lambda;        /* final damping */
  int32_t n_opt_cams;
  int32_t gpu_launches;
} hfb_lba_stats;

/* optimizer.optimize(iterations) (src/Optimizer.cc:1411; Thirdparty/g2o/g2o/core/sparse_optimizer.cpp:353-416,
 * optimization_algorithm_levenberg.cpp:61-169).  stop_flag mirrors `bool* pbStopFlag` (src/Optimizer.cc:1203-1204),
 * may be NULL.  user_lambda_init <= 0: tau * max diagonal (levenberg.cpp:171-185).
 * Outputs: poses_out [n_cams*7], points_out [n_points*3], chi2_out [n_edges] (as cached by the last error evaluation,
 * the value `e->chi2()` returns at src/Optimizer.cc:1425), depth_positive_out [n_edges] (isDepthPositive()). */
int hfb_lba_optimize(hfb_ctx* ctx, const hfb_lba_problem* problem, int32_t iterations, double user_lambda_init,
                     const volatile uint8_t* stop_flag, double* poses_out, double* points_out, double* chi2_out,
                     uint8_t* depth_positive_out, hfb_lba_stats* stats);

/* Optimizer::PoseOptimization (src/Optimizer.cc:814-1114), monocular branch: motion-only BA of one frame against fixed
 * map points.  4 rounds x optimize(10) (Levenberg, Huber sqrt(5.991) dropped after the third round), inlier
 * re-classification with chi2 > 5.991 after every round, every round restarted from pose_in -- the whole schedule
 * runs in one device kernel.  pose: qx qy qz qw tx ty tz (Tcw).  outlier_out = Frame::mvbOutlier of the used edges,
 * *n_inliers = nInitialCorrespondences - nBad. */
int hfb_pose_optimize(hfb_ctx* ctx, const float* K, const double* pose_in, int32_t n, const double* Xw,
                      const double* obs, const double* inv_sigma2, double* pose_out, uint8_t* outlier_out,
                      int32_t* n_inliers, int32_t* n_trials);

/* One linearisation + Schur reduction at the given estimate and damping (parity hook for the two kernels):
 * Hschur [6n_opt x 6n_opt] row-major (full symmetric), bschur [6n_opt], robust chi2 (sum of rho(chi2_e)).
 * Optimisable cameras take slots in array order. */
int hfb_lba_build_schur(hfb_ctx* ctx, const hfb_lba_problem* problem, double lambda, double* Hschur, double* bschur,
                        double* robust_chi2, int32_t* n_opt_cams);

#ifdef __cplusplus
}
#endif
#endif /* HFNET_B200_H_ */
