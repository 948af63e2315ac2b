/*
 * brtpe.h -- C ABI of libbrtpe.so: hand-written sm_100a CUDA kernels for the
 * HigherHRNet-W48 teacher-inference hot path of andres-fr/realtime-pose-estimation.
 *
 * Every entry point replaces a piece of the reference's Python hot path (file:line
 * are into /root/reference).  Conventions:
 *   - plain pointers and sizes only; all data pointers are DEVICE pointers unless a
 *     parameter is documented as host memory;
 *   - every function returns 0 on success or a negative BRTPE_E* code;
 *     brtpe_last_error() returns a thread-local, human readable message;
 *   - no hidden device allocation: scratch comes from the caller
 *     (*_workspace_bytes() queries), nothing is retained after the call except by the
 *     explicit plan objects (brtpe_plan_*);
 *   - kernels are enqueued on `stream` (a cudaStream_t passed as void*), no host
 *     synchronisation inside unless documented.
 */
#ifndef BRTPE_H_
#define BRTPE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BRTPE_OK 0
#define BRTPE_EINVAL (-1)      /* bad argument / unsupported configuration           */
#define BRTPE_ECUDA (-2)       /* a CUDA runtime / driver call failed                */
#define BRTPE_EWORKSPACE (-3)  /* workspace too small                                */
#define BRTPE_EOVERFLOW (-4)   /* a caller-sized output capacity was exceeded        */

#define BRTPE_MAX_TAG_DIMS 4   /* T: tag copies per joint (1 = no flip, 2 = flip)    */
#define BRTPE_MAX_TOPK 64      /* K = max_num_people for top-k                        */
#define BRTPE_MAX_GROUP_K 32   /* K supported by the grouping (Hungarian) kernel      */
#define BRTPE_MAX_JOINTS 32

const char* brtpe_last_error(void);
int brtpe_version(void);
/* 1 if the library was built with the tcgen05 convolution path */
int brtpe_has_umma(void);

/* ------------------------------------------------------------------------------------
 * Decode: HeatmapParser (rtpe/third_party/group.py:125-287)
 * ---------------------------------------------------------------------------------- */

/* Mirror of Params (group.py:100-122) + HeatmapParser ctor kwargs (group.py:126-132). */
typedef struct brtpe_decode_params {
  int32_t num_joints;            /* J                                                  */
  int32_t max_num_people;        /* K                                                  */
  double detection_threshold;    /* compared in float64 like the reference (group.py:41) */
  double tag_threshold;          /* group.py:85                                        */
  int32_t use_detection_val;
  int32_t ignore_too_much;
  int32_t tag_per_joint;         /* 1: tag has J planes, 0: one plane shared by joints */
  int32_t nms_ksize;             /* odd; 2*nms_padding must equal nms_ksize-1          */
  int32_t nms_padding;
  int32_t munkres_start_rule;    /* 0: munkres 1.1.x (restart at previous hit), 1: <=1.0.x */
} brtpe_decode_params;

/* HeatmapParser.nms (group.py:134-138): out = det * (maxpool_{k,1,p}(det) == det).
 * det, out: (N*J, H, W) float32. */
int brtpe_nms(const float* det, float* out, int planes, int H, int W, int ksize,
              int padding, void* stream);

/* HeatmapParser.top_k (group.py:144-179), fused NMS + per-(image, joint) top-K +
 * tag gather.  Tie rule: value descending, flat index ascending.
 *   det   (N, J, H, W) f32          tag (N, Jt, H, W, T) f32, Jt = J or 1
 *   val_k (N, J, K) f32             ind_k (N, J, K) i32 (flat index y*W+x)
 *   loc_k (N, J, K, 2) i64 (x, y)   tag_k (N, J, K, T) f32
 * loc_k may be NULL. */
size_t brtpe_topk_workspace_bytes(int N, int J, int H, int W, int K);
int brtpe_nms_topk_gather(const float* det, const float* tag, int N, int J, int Jt, int H,
                          int W, int T, int K, int ksize, int padding, float* val_k,
                          int32_t* ind_k, int64_t* loc_k, float* tag_k, void* workspace,
                          size_t workspace_bytes, void* stream);

/* HeatmapParser.match / match_by_tag / py_max_match (group.py:19-97, :140-142):
 * per-image greedy-by-joint grouping with an exact Kuhn-Munkres assignment per joint
 * round (float64 cost, munkres 1.1.x tie-breaking), one warp per image.
 *   val_k (N,J,K) f32, ind_k (N,J,K) i32, tag_k (N,J,K,T) f32, W = map width
 *   ans   (N, Pmax, J, 3+T) f32  persons in creation order: [x, y, val, tag...]
 *   count (N) i32                number of persons per image (<= Pmax or EOVERFLOW flag)
 * Requires K <= BRTPE_MAX_GROUP_K, T <= BRTPE_MAX_TAG_DIMS.  Pmax = J*K always fits.
 * overflow (1) i32 device flag is set to 1 if any image needed more than Pmax. */
size_t brtpe_group_workspace_bytes(int N, int J, int K, int T, int Pmax);
int brtpe_group_ae(const float* val_k, const int32_t* ind_k, const float* tag_k, int N,
                   int W, int T, const brtpe_decode_params* params, float* ans,
                   int32_t* count, int32_t* overflow, int Pmax, void* workspace,
                   size_t workspace_bytes, void* stream);

/* HeatmapParser.adjust (group.py:181-200): quarter-pixel shift + 0.5, in place on ans.
 * det (N,J,H,W) f32 is the un-NMS'd map given to parse(). */
int brtpe_adjust(float* ans, const int32_t* count, const float* det, int N, int J, int H,
                 int W, int T, int Pmax, void* stream);

/* scores = mean over joints of ans[..., 2] with numpy's float32 pairwise order
 * (group.py:272).  scores (N, Pmax) f32. */
int brtpe_scores(const float* ans, const int32_t* count, float* scores, int N, int J, int T,
                 int Pmax, void* stream);

/* HeatmapParser.refine (group.py:202-264) for every person of every image: one pass
 * over det + tag per image evaluates argmax(det_j - round(||tag_j - mean_tag||)) for the
 * joints a person is missing and fills them in place.
 *   det (N,J,H,W) f32, tag (N,Jt,H,W,T) f32, ans (N,Pmax,J,3+T) f32 in/out. */
size_t brtpe_refine_workspace_bytes(int N, int J, int T, int Pmax);
int brtpe_refine(const float* det, const float* tag, float* ans, const int32_t* count, int N,
                 int J, int Jt, int H, int W, int T, int Pmax, void* workspace,
                 size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------
 * Aggregation between network and parser
 * ---------------------------------------------------------------------------------- */

/* F.interpolate(src, (Ho,Wo), mode="bilinear", align_corners=ac) on NCHW float32
 * (validate_hhrnet.py:94-98).  dst plane p / channel stride given in elements so the
 * result can land inside a (N, J, Ho, Wo, T) tag tensor: dst[(p*Ho*Wo + y*Wo + x)*dst_inner + dst_off].
 * src (planes, Hi, Wi) with plane stride src_plane_stride elements. */
int brtpe_bilinear_resize(const float* src, long long src_plane_stride, int planes, int Hi,
                          int Wi, float* dst, int Ho, int Wo, int dst_inner, int dst_off,
                          int align_corners, void* stream);

/* Image pre-processing of the teacher-inference driver, fused:
 *   cv2.warpAffine(image, trans, (Wo, Ho))  -- rtpe/third_party/transforms.py:183-190 (uint8 HxWx3,
 *       INTER_LINEAR, BORDER_CONSTANT 0; OpenCV's fixed-point arithmetic reproduced bit for bit)
 *   torchvision ToTensor() + Normalize(mean, std) -- teacher_inference.py:70-73, :79
 * img: device uint8 (Hs, Ws, 3) with row pitch src_pitch bytes; trans_host: host double[6], the
 * forward 2x3 matrix exactly as the reference hands it to cv2.warpAffine; mean_host / std_host:
 * host float[3].  out_f32: device (3, Ho, Wo) float32 = ((u8 / 255) - mean) / std, or NULL;
 * out_u8: device (Ho, Wo, 3) uint8 = the warped image, or NULL (at least one is required). */
int brtpe_preprocess_warp_normalize(const uint8_t* img, int Hs, int Ws, int src_pitch,
                                    const double* trans_host, int Ho, int Wo,
                                    const float* mean_host, const float* std_host, float* out_f32,
                                    uint8_t* out_u8, void* stream);

/* Flip-test / multi-scale aggregation of ONE scale (upstream HigherHRNet
 * get_multi_stage_outputs + aggregate_results; in-tree callers legacy/valid_ae_avg.py:176-185,
 * legacy/valid_ae1dim.py:176-191; configuration legacy/distillation.py:85-92):
 *   y0 (N, J+A, H4, W4), y1 (N, J, H2, W2): network outputs of the image;
 *   y0f, y1f: outputs of the x-mirrored image, or NULL (no flip test).
 *   stage(y0,y1)[c] = ( bilinear_{H4->H2}(y0[c]) + y1[c] ) / 2          (align_corners=False)
 *   heat = bilinear_{H2->Hb}(stage(y0,y1)); with flip: heat = (heat + heat_f) / 2 where
 *          heat_f is the same on (y0f,y1f), mirrored in x and channel-permuted by flip_index
 *   det (N, J, Hb, Wb) = accumulate ? det + heat : heat, then / final_div if final_div != 0
 *   tag_out (N, A, Hb, Wb, T), T = 1 + (flip): slot 0 = bilinear_{H2->Hb}(bilinear_{H4->H2}(y0[J+a])),
 *          slot 1 = the mirrored / permuted flipped-image tags.  tag_out may be NULL.
 * flip_index_host: host int32[J] (also used for the A tag channels when A == J). */
int brtpe_aggregate_scale(const float* y0, const float* y1, const float* y0f, const float* y1f,
                          int N, int J, int A, int H4, int W4, int H2, int W2, int Hb, int Wb,
                          const int32_t* flip_index_host, int accumulate, float final_div,
                          float* det, float* tag_out, void* stream);

/* ------------------------------------------------------------------------------------
 * Convolution engine: PoseHigherResolutionNet.forward (pose_higher_hrnet.py:637-686)
 * ---------------------------------------------------------------------------------- */

/* BRTPE_DT_BF16X2 ("split fp32", the float32 mode of the tcgen05 engines): every activation value v
 * is stored as TWO bf16 numbers hi = bf16(v), lo = bf16(v - hi) (16 significant bits together).  A
 * pixel of stride ld bf16 elements holds the hi parts in channels [0, ld/2) and the lo parts in
 * [ld/2, ld): all channel offsets / counts of a descriptor index the hi half, the lo part of channel
 * c sits at c + ld/2.  Convolutions accumulate hi*w_hi + lo*w_hi + hi*w_lo in float32 (three bf16
 * MMAs per product, relative error ~1e-5 against float32: fp32 mode <= 1e-4). */
enum { BRTPE_DT_F32 = 0, BRTPE_DT_BF16 = 1, BRTPE_DT_BF16X2 = 2 };
/* FFMA: CUDA-core fp32-accurate path.  UMMA: tcgen05, one TMA box per tap (any tap table).
 * UMMA_HALO: tcgen05, 3x3/stride-1 only, one halo tile per channel block + cluster-multicast
 * weights.  AUTO picks HALO, then UMMA, then FFMA, by what the layer allows. */
enum { BRTPE_ENGINE_AUTO = 0, BRTPE_ENGINE_FFMA = 1, BRTPE_ENGINE_UMMA = 2,
       BRTPE_ENGINE_UMMA_HALO = 3 };

/* One fused conv launch: out = act( conv(in, w) + bias [+ residual] ), NHWC activations.
 * Covers nn.Conv2d 3x3/1x1 stride 1/2 (pose_higher_hrnet.py:40-43,:83-89,:163,:202,:218,
 * :460-482,:558-580) with eval BatchNorm folded into (w, bias), the ReLU / residual adds
 * of BasicBlock/Bottleneck (:57-75,:96-116), and one output-parity phase of
 * ConvTranspose2d(k4,s2,p1) (:514-521) expressed through the tap table. */
typedef struct brtpe_conv_desc {
  int32_t dtype;          /* BRTPE_DT_*: activation + weight storage type              */
  int32_t engine;         /* BRTPE_ENGINE_*                                            */
  int32_t N, Hin, Win;    /* input  tensor (N, Hin, Win, .) NHWC                        */
  int32_t Cin;            /* logical input channels (K per tap)                         */
  int32_t in_ld;          /* input pixel stride in elements (>= in_coff + Cin)          */
  int32_t in_coff;        /* first input channel inside the pixel                       */
  int32_t Hm, Wm;         /* GEMM-M domain (output pixels computed): m = (n, ym, xm)    */
  int32_t in_stride;      /* input pixel = (ym*in_stride + dy, xm*in_stride + dx)       */
  int32_t ntaps;          /* <= 9                                                       */
  int32_t tap_dy[9], tap_dx[9];
  int32_t Hout, Wout;     /* output tensor (N, Hout, Wout, .)                           */
  int32_t out_scale;      /* output pixel = (ym*out_scale + out_oy, xm*out_scale + out_ox) */
  int32_t out_oy, out_ox;
  int32_t Cout;           /* logical output channels                                    */
  int32_t out_ld, out_coff;
  int32_t res_ld, res_coff;   /* residual tensor has the output's (N,Hout,Wout) geometry */
  int32_t relu;
  int32_t Cout_store;     /* channels written per pixel (>= Cout; extra ones get 0)     */
  /* HRNet cross-resolution fuse-add in the epilogue (HighResolutionModule.forward,
   * pose_higher_hrnet.py:245-254: y_i = relu(sum_j f_ij(x_j))): up to three extra addends, term k a
   * (N, Hout >> add_shift[k], Wout >> add_shift[k], add_ld[k]) tensor of Cout channels read with
   * nearest-neighbour upsampling (nn.Upsample(scale_factor=2^shift), :202-209).
   *   out2_ld == 0:  out  = act(conv + bias + residual + sum_k up(add_k))
   *   out2_ld  > 0:  out  = act(conv + bias + residual)            (the branch output x_i)
   *                  out2 = relu(out + sum_k up(add_k))             (the fused output y_i)
   * tcgen05 engines only, whole 16-channel chunks (Cout == Cout_store, a multiple of 16). */
  int32_t n_add;
  int32_t add_ld[3], add_shift[3];
  int32_t out2_ld;
  /* 1: walk the output tiles from the last image to the first (halo engine).  Activations at the
   * bench's chunk are larger than the 126 MB L2: a layer that starts where its producer ended finds the
   * most recently written / read part of its input and residual still in the L2. */
  int32_t reverse_order;
} brtpe_conv_desc;

/* FFMA path weights: float32 [ntaps][Cin][Cout].  UMMA path weights: bf16
 * [ntaps][Cout_pad][Cin_pad] (K-major), Cin_pad = roundup(Cin, 64), Cout_pad = roundup(Cout,16).
 * bias: float32 [Cout] (may be NULL).  residual may be NULL. */
int brtpe_conv_run(const brtpe_conv_desc* d, const void* in, const void* weights,
                   const float* bias, const void* residual, void* out, void* stream);

/* brtpe_conv_run for a descriptor with fuse addends: add_ptrs[n_add] device tensors, out2 the second
 * output (NULL unless out2_ld > 0). */
int brtpe_conv_run_fused(const brtpe_conv_desc* d, const void* in, const void* weights,
                         const float* bias, const void* residual, void* out,
                         const void* const* add_ptrs, void* out2, void* stream);

/* BasicBlock body in ONE launch (rtpe/third_party/pose_higher_hrnet.py:46-75: conv1 + bn1 + relu ->
 * conv2 + bn2, += residual, relu): d0 / d1 describe two 3x3 stride-1 bf16 layers of equal geometry and
 * width (<= 64 channels, both with ReLU), `mid` receives conv1's output (written in full, as by two
 * brtpe_conv_run calls), `residual` is added by conv2.  The launch walks the batch in sub-batches with
 * conv2 one phase behind conv1, so that a sub-batch's input, intermediate and output are still in the
 * L2 when conv2 reads them (three tensor passes through HBM instead of five at batch sizes whose
 * tensors exceed the L2).  Results are bit-identical to two brtpe_conv_run calls.  Plans use it on
 * their own (BRTPE_CHAIN=0 switches it off); BRTPE_EINVAL when the pair is not chainable. */
int brtpe_conv_chain_run(const brtpe_conv_desc* d0, const void* in, const void* w0,
                         const float* bias0, void* mid, const brtpe_conv_desc* d1, const void* w1,
                         const float* bias1, const void* residual, void* out, void* stream);

/* Engine brtpe_conv_run / brtpe_plan_add_conv will use for this descriptor
 * (BRTPE_ENGINE_FFMA, BRTPE_ENGINE_UMMA or BRTPE_ENGINE_UMMA_HALO), or a negative error code. */
int brtpe_conv_select_engine(const brtpe_conv_desc* d);
/* Packed-weight geometry of the tcgen05 path: bf16 [ntaps][*cout_pad][*cin_pad]. */
int brtpe_umma_weight_dims(int Cin, int Cout_store, int* cin_pad, int* cout_pad);

/* ---- weight pre-packing: eval BatchNorm folded into its convolution + engine layout, one launch
 * per layer (pose_higher_hrnet.py conv + bn pairs, e.g. :57-64, :96-106, :202-229, :514-521;
 * replaces nn.BatchNorm2d.forward in eval mode).  float32 arithmetic, identical to
 *   scale = gamma / sqrt(var + eps);  w' = w * scale[co];  b' = beta - mean * scale (+ bias * scale)
 * evaluated with separate IEEE operations. */
enum { BRTPE_WT_F32 = 0, BRTPE_WT_BF16 = 1, BRTPE_WT_F16 = 2 };
enum { BRTPE_PACK_CIN_COUT_F32 = 0,   /* float32 [ntaps][Cin_store][Cout_pack] (CUDA-core engine)   */
       BRTPE_PACK_KMAJOR_BF16 = 1 };  /* bf16 [ntaps][cout_pad][cin_pad] (tcgen05 engines)           */
typedef struct brtpe_prepack_desc {
  int32_t w_dtype;          /* BRTPE_WT_*: storage type of w and conv_bias                       */
  int32_t transposed;       /* 0: w (Cout, Cin, KH, KW) nn.Conv2d; 1: w (Cin, Cout, KH, KW)
                               nn.ConvTranspose2d                                                */
  int32_t Cout, Cin, KH, KW;
  int32_t ntaps;            /* packed taps (<= 9)                                                */
  int32_t tap_kh[9], tap_kw[9];   /* kernel position packed tap t takes its weights from         */
  int32_t im2col;           /* 1 (ntaps == 1): stored input channel k = (kh*KW + kw)*Cin + c, the
                               operand of brtpe_stem_im2col                                      */
  int32_t Cin_store;        /* stored input channels per tap (K); channels >= Cin get zeros      */
  int32_t layout;           /* BRTPE_PACK_*                                                      */
  int32_t Cout_pack;        /* layout 0: row length (>= Cout; the extra columns are zero)        */
  int32_t cin_pad, cout_pad;      /* layout 1: brtpe_umma_weight_dims                            */
  int32_t round_bf16;       /* layout 0: values rounded to bf16 (the operand rounding of tcgen05) */
  float bn_eps;
  int32_t split;            /* layout 1, BRTPE_DT_BF16X2 layers: K = three segments of
                               roundup(Cin_store, 64) channels [w_hi | w_hi | w_lo] (cin_pad = 3x),
                               w_hi = bf16(w'), w_lo = bf16(w' - w_hi)                            */
} brtpe_prepack_desc;
/* w, conv_bias (NULL: none): device, d->w_dtype.  bn_* (all NULL: no BatchNorm): device float32
 * [Cout].  cin_index (NULL: identity): device int32 [Cin_store], module input channel carried by
 * stored channel k, or -1 for a zero (pad) channel.  packed: device, every element is written.
 * bias_out (may be NULL): device float32 [bias_len], entries >= Cout are zero. */
int brtpe_prepack_weights(const brtpe_prepack_desc* d, const void* w, const void* conv_bias,
                          const float* bn_gamma, const float* bn_beta, const float* bn_mean,
                          const float* bn_var, const int32_t* cin_index, void* packed,
                          float* bias_out, int bias_len, void* stream);

/* Stem conv1 (pose_higher_hrnet.py:363-365,:638-640): NCHW float32/half image ->
 * 3x3 s2 conv(3->Cout) + folded BN + ReLU -> NHWC (f32 or bf16).  w: float32 [27][Cout]
 * ordered (ky, kx, cin). */
int brtpe_stem_conv1(const void* img, int img_dtype_is_half, int N, int H, int W,
                     const float* w, const float* bias, int Cout, void* out, int out_dtype,
                     void* stream);

/* Stem im2col, the bf16 / tcgen05 form of conv1 (pose_higher_hrnet.py:363-365): NCHW
 * float32/half image -> NHWC bf16 (N, H/2, W/2, 32); channel k = (ky*3 + kx)*3 + ci is the
 * 3x3 / stride-2 / pad-1 window of output pixel (oy, ox), channels 27..31 are zero.  conv1
 * + BN + ReLU is then brtpe_conv_run with a 1x1 tap table and Cin = 32.
 * img_mode: bit 0 = the image is half (else float32); bit 1 = flip pair: img holds N/2 images and
 * image n >= N/2 of the batch is image n - N/2 mirrored in x (the flip-test batch cat(x, flip(x))
 * of the upstream get_multi_stage_outputs caller, never materialised); bit 2 = a float32 image is
 * rounded through fp16 first, like fp16util.py's tofp16 in front of the network. */
int brtpe_stem_im2col(const void* img, int img_mode, int N, int H, int W, void* out,
                      void* stream);

/* y_i = relu?( sum_k up_{2^shift_k}(term_k) ) of HighResolutionModule.forward
 * (pose_higher_hrnet.py:245-254); NHWC, nterms <= 4, nearest-neighbour upsample. */
int brtpe_fuse_sum(int dtype, int nterms, const void* const* terms, const int32_t* shifts,
                   const int32_t* term_ld, int N, int H, int W, int C, void* out, int out_ld,
                   int relu, void* stream);

/* NHWC (f32|bf16, pixel stride ld, channel offset coff) -> NCHW float32 or half
 * network outputs (tofp32, fp16util.py:54-68). */
int brtpe_nhwc_to_nchw(int dtype, const void* src, int N, int H, int W, int C, int ld,
                       int coff, void* dst, int dst_is_half, void* stream);

/* ---- execution plans: a recorded list of launches replayed natively / as a CUDA graph */
typedef struct brtpe_plan brtpe_plan;
brtpe_plan* brtpe_plan_create(void);
void brtpe_plan_destroy(brtpe_plan*);
int brtpe_plan_add_conv(brtpe_plan*, const brtpe_conv_desc* d, const void* in,
                        const void* weights, const float* bias, const void* residual,
                        void* out);
/* addend tensors / second output of the conv op added last (descriptor with n_add > 0) */
int brtpe_plan_set_conv_fuse(brtpe_plan*, const void* const* add_ptrs, void* out2);
int brtpe_plan_add_stem(brtpe_plan*, const void* img, int img_is_half, int N, int H, int W,
                        const float* w, const float* bias, int Cout, void* out, int out_dtype);
int brtpe_plan_add_fuse(brtpe_plan*, int dtype, int nterms, const void* const* terms,
                        const int32_t* shifts, const int32_t* term_ld, int N, int H, int W,
                        int C, void* out, int out_ld, int relu);
int brtpe_plan_add_nhwc_to_nchw(brtpe_plan*, int dtype, const void* src, int N, int H, int W,
                                int C, int ld, int coff, void* dst, int dst_is_half);
int brtpe_plan_add_stem_im2col(brtpe_plan*, const void* img, int img_mode, int N, int H, int W,
                               void* out);
/* Scheduling annotation of the op added last: `lane` (0..7) is the capture stream the op is
 * issued on, deps[ndeps] are indices of EARLIER ops that must have completed before it (ops
 * of the same lane are ordered implicitly).  brtpe_plan_graph_launch turns lanes into
 * parallel branches of the CUDA graph (the HRNet resolution branches run concurrently);
 * brtpe_plan_run / brtpe_plan_profile ignore the annotation and run in recording order.
 * Ops without annotation are lane 0 with no cross-lane dependencies. */
int brtpe_plan_set_sched(brtpe_plan*, int lane, const int32_t* deps, int ndeps);
int brtpe_plan_num_ops(const brtpe_plan*);
/* total algorithmic conv FLOPs (2*MAC) of the plan's conv ops */
double brtpe_plan_conv_flops(const brtpe_plan*);
/* enqueue every op on `stream` */
int brtpe_plan_run(brtpe_plan*, void* stream);
/* capture the op list into a CUDA graph (once), then launch it */
int brtpe_plan_graph_launch(brtpe_plan*, void* stream);
/* run un-graphed with CUDA events around each op; ms_out[num_ops] (host) gets per-op
 * device milliseconds; kinds_out[num_ops] (host): 0 conv-umma, 1 conv-ffma, 2 other, 3 conv-umma-halo.
 * Synchronises the stream. */
int brtpe_plan_profile(brtpe_plan*, void* stream, float* ms_out, int32_t* kinds_out,
                       double* flops_out);

/* ---- context-aware-module student (rtpe/students.py:118-201, :595-771; BASELINE config 4)
 * Small NHWC streaming ops, dtype = BRTPE_DT_* of the activations.  iparams[0] is always the
 * dtype; `kind` and the remaining parameters:
 *  1 AvgPool2d(3,2,1,count_include_pad=False) (students.py:657): in0 (N,H,W,in_ld) -> out
 *    (N,H/2,W/2,out_ld);                       iparams = {dtype, N, H, W, C, in_ld, out_ld}
 *  2 SELayer pooling, stage 1 (students.py:137-139): in0 (N,HW,ld) -> out float
 *    partial[N][chunks][C];                    iparams = {dtype, N, HW, C, ld, chunks}
 *  3 SELayer gate (students.py:129-141): in0 = partial, in1 = float [W1 (hid x C), b1 (hid),
 *    W2 (C x hid), b2 (C)] -> out float gate[N][C] = sigmoid(W2 relu(W1 mean + b1) + b2);
 *                                              iparams = {dtype, N, C, hid, chunks, HW}
 *  4 ContextAwareModule tail (students.py:199-200): out = relu(in0 + in1 * gate), in2 = gate;
 *                                              iparams = {dtype, N, HW, C, ld0, ld1, ld_out}
 *  5 attention injection (students.py:752-753): a = sigmoid(in0[...,0] / 20); out = in1 + a;
 *    in2 = float att_out (N,H,W) written with a; iparams = {dtype, N, HW, C, ld_att, ld_in1, ld_out}
 *  6 NHWC bilinear resize (MultistageStudent's out_hw, students.py:481-498): in0 (N,H,W,in_ld) ->
 *    channels [coff, coff + C) of out (N,Ho,Wo,out_ld);
 *                                  iparams = {dtype, N, H, W, C, in_ld, out_ld, Ho, Wo, align_corners, coff}
 *  7 NCHW image -> NHWC slice (AttentionStudentSteps' alt image, students.py:980-988): in0 = float
 *    (or half) image (N,C,H,W); s2d = 0: bilinear (align_corners=False) resize to (Ho,Wo) into channels
 *    [coff, coff + C), zeros in [coff + C, coff + Cz); s2d = 1: space-to-depth by 2, channel
 *    (ry*2+rx)*C + c of pixel (y,x) = img[c][2y+ry][2x+rx], zeros in [4C, Cz);
 *                                  iparams = {dtype, N, H, W, C, img_is_half, out_ld, Ho, Wo, coff, Cz, s2d}
 *  8 NHWC space-to-depth by 2: in0 (N,H,W,in_ld) -> out (N,H/2,W/2,out_ld), channel (ry*2+rx)*C + c;
 *    with it a 5x5 / stride-2 conv (students.py:835-846) is a 3x3 / stride-1 conv over 4C channels;
 *                                  iparams = {dtype, N, H, W, C, in_ld, out_ld}
 *  9 attention product (students.py:1001-1018): a = sigmoid(in0[...,0] / div); out = in1 * a; in2 =
 *    float att_out (N,H,W) written with a;
 *                                  iparams = {dtype, N, HW, C, ld_att, ld_in1, ld_out, float bits of div}
 * Kinds 3 and 4 take optional trailing parameters: 3 {..., Cin} = number of pooled (stored) channels
 * when it differs from the gate's C; 4 {..., Cz} = zero-fill channels [C, Cz) of out.
 * nparams <= 12.
 */
int brtpe_aux_run(int kind, const void* in0, const void* in1, const void* in2, void* out,
                  const int32_t* iparams, int nparams, void* stream);
int brtpe_plan_add_aux(brtpe_plan*, int kind, const void* in0, const void* in1, const void* in2,
                       void* out, const int32_t* iparams, int nparams);

/* ---- debug instrumentation (not part of the reference surface)
 * While `buf` is non-NULL every conv_halo_kernel launch with <= max_ctas CTAs runs its
 * instrumented instantiation and writes 16 int64 cycle counters per CTA to
 * buf[cta*16 + slot] (device memory, caller owned; slots documented in csrc/conv_halo.cu).
 * Pass NULL to switch it off. */
int brtpe_debug_halo_prof(void* buf, int max_ctas);

#ifdef __cplusplus
}
#endif
#endif /* BRTPE_H_ */
