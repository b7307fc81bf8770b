/*
 * saf_oracle.c -- CPU restatement of the reference's RGB-D fusion hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in spatially_aware_ai_b200/ may import, link or call
 * this file; it is the checker used by tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py.
 *
 * Parity status: PINNED by running the unmodified reference (torch-CPU) in the build
 * container and committing its inputs/outputs as tests/golden/(all).npz (generator:
 * tests/golden/make_golden.py).  The reference's own repository holds no golden vectors or
 * automated tests for this path.
 *
 * What is restated (all line numbers in /root/reference):
 *   voxel centres          clip_seem_fusion.py:664-672   (clipfusion.py:617-625)
 *   projection             clip_seem_fusion.py:698-712   (clipfusion.py:648-659)
 *   nearest depth sample   clip_seem_fusion.py:714-719   (clipfusion.py:661-666)
 *   sdf, masks             clip_seem_fusion.py:722-728   (clipfusion.py:669-679)
 *   tsdf / tsdf_weight     clip_seem_fusion.py:730-744   (clipfusion.py:681-695)
 *   label / rgb / feature  clip_seem_fusion.py:786-805   (clipfusion.py:701-713)
 *   running averages       clip_seem_fusion.py:808-814   (clipfusion.py:715-721)
 *   label histogram        clip_seem_fusion.py:820-822
 * The arithmetic the reference delegates to torch (third-party, pinned torch==2.0.0+cu118 in
 * environment.yml:251; executed here under torch 2.11.0 CPU) is restated from its published
 * semantics: bmm with K=3 is the left-to-right fma chain, every other op is one correctly
 * rounded fp32 operation in source order, grid_sample follows ATen
 * native/GridSampler.h (unnormalize, align_corners=False) and
 * native/cpu/GridSamplerKernel.cpp (nearest = round-half-even, zeros padding, bilinear
 * weights (1-w)(1-n).. and tap order nw,ne,sw,se).
 *
 * Build: gcc -O2 -ffp-contract=off (contraction would change roundings), optional -fopenmp.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define SAF_ORACLE_MAX_BATCH 8

/* ATen grid_sampler_unnormalize, align_corners=False: ((g+1)*size - 1)/2 (GridSampler.h:27-36).
 * Both ATen builds evaluate it with ONE rounding after the add: the CPU kernel is written
 * (g+1)*(size/2) - 0.5 and is compiled with fp contraction, the CUDA kernel's (g+1)*size-1 is
 * contracted by nvcc and the halving is exact.  Verified bit-for-bit against torch-CPU
 * bilinear outputs; for nearest the two forms pick the same pixel (exhaustive scans). */
static inline float unnormalize(float g, int size)
{
    return fmaf(g + 1.0f, (float)size / 2.0f, -0.5f);
}

/* nearest tap index or -1 when out of bounds / NaN (zeros padding). */
static inline int nearest_index(float g, int size)
{
    float r = nearbyintf(unnormalize(g, size)); /* half-to-even, default rounding mode */
    if (!(r >= 0.0f && r < (float)size))
        return -1;
    return (int)r;
}

typedef struct {
    int idx[4];   /* flat pixel index y*W+x or -1 (dropped tap), order nw ne sw se */
    float w[4];
} bilinear_taps;

/* ATen CPU bilinear interpolation parameters (GridSamplerKernel.cpp, compute_interp_params). */
static inline void bilinear_setup(float gx, float gy, int W, int H, bilinear_taps *t)
{
    float x = unnormalize(gx, W);
    float y = unnormalize(gy, H);
    float x_w = floorf(x), y_n = floorf(y);
    float w = x - x_w, e = 1.0f - w;
    float n = y - y_n, s = 1.0f - n;
    t->w[0] = s * e;
    t->w[1] = s * w;
    t->w[2] = n * e;
    t->w[3] = n * w;
    /* NaN / huge coordinates behave as out of bounds (cvttps -> INT_MIN in ATen). */
    int ok = (x_w >= -2.0f && x_w <= (float)W + 1.0f && y_n >= -2.0f && y_n <= (float)H + 1.0f);
    int ixw = ok ? (int)x_w : -2, iyn = ok ? (int)y_n : -2;
    int ixe = ixw + 1, iys = iyn + 1;
    int wm = ixw > -1 && ixw < W, em = ixe > -1 && ixe < W;
    int nm = iyn > -1 && iyn < H, sm = iys > -1 && iys < H;
    t->idx[0] = (nm && wm) ? iyn * W + ixw : -1;
    t->idx[1] = (nm && em) ? iyn * W + ixe : -1;
    t->idx[2] = (sm && wm) ? iys * W + ixw : -1;
    t->idx[3] = (sm && em) ? iys * W + ixe : -1;
}

static inline float bilinear_eval(const bilinear_taps *t, const float *img, ptrdiff_t pix_stride)
{
    float v0 = t->idx[0] >= 0 ? img[(ptrdiff_t)t->idx[0] * pix_stride] : 0.0f;
    float v1 = t->idx[1] >= 0 ? img[(ptrdiff_t)t->idx[1] * pix_stride] : 0.0f;
    float v2 = t->idx[2] >= 0 ? img[(ptrdiff_t)t->idx[2] * pix_stride] : 0.0f;
    float v3 = t->idx[3] >= 0 ? img[(ptrdiff_t)t->idx[3] * pix_stride] : 0.0f;
    /* ATen CPU: nw*w0 + ne*w1 + sw*w2 + se*w3, compiled to a left-to-right fma chain */
    return fmaf(v3, t->w[3], fmaf(v2, t->w[2], fmaf(v1, t->w[1], v0 * t->w[0])));
}

/* clip_seem_fusion.py:698-712 : voxel centre -> normalised image coords (gx, gy) and depth z */
static inline void project(const float *pose, const float *K, float xw, float yw, float zw,
                           int W, int H, float *gx, float *gy, float *z)
{
    float d0 = xw - pose[3], d1 = yw - pose[7], d2 = zw - pose[11];
    float xc[3], uvz[3];
    for (int k = 0; k < 3; ++k) /* row k of R^T = column k of R */
        xc[k] = fmaf(pose[8 + k], d2, fmaf(pose[4 + k], d1, pose[k] * d0));
    for (int k = 0; k < 3; ++k)
        uvz[k] = fmaf(K[3 * k + 2], xc[2], fmaf(K[3 * k + 1], xc[1], K[3 * k] * xc[0]));
    float u = uvz[0] / uvz[2];
    float v = uvz[1] / uvz[2];
    *z = uvz[2];
    *gx = ((u + 0.5f) / (float)W) * 2.0f - 1.0f;
    *gy = ((v + 0.5f) / (float)H) * 2.0f - 1.0f;
}

/*
 * One reference integrate() call on an x-slab [x_begin, x_end) of the grid.
 * State arrays are slab-local (voxel 0 = (x_begin,0,0)); coordinates use the global x index.
 *
 *   depth [B,H,W] f32; rgb [B,H,W,3] f32; seg [B,H,W] f32 class ids or NULL (ClipFusion);
 *   table: feature image of frame b at table + b*table_sb, element (c, r) at c*table_sc + r*table_sr
 *          with r = py*npx + px  (reference layout [C,npy,npx]: sc = npy*npx, sr = 1);
 *   poses [B,16] cam->world row-major; K [B,9];
 *   rgb_mode 0 = nearest (clipfusion.py:701-706), 1 = bilinear (clip_seem_fusion.py:793-798);
 *   labels [Nslab, n_classes] or NULL;
 *   valid_out / tsdf_valid_out: optional [B,Nslab] u8 masks; counts: optional [2*B] (valid, tsdf_valid per frame).
 * Returns 0, or a negative code for bad arguments, or 1 if a class id was outside [0, n_classes)
 * (the reference's one_hot raises there).
 */
int saf_oracle_integrate_ex(const float *origin, float voxel_size, const int *nvox, int x_begin, int x_end,
                            float trunc, int B, int H, int W, const float *depth, const float *rgb,
                            const float *seg, const float *table, ptrdiff_t table_sb, ptrdiff_t table_sc,
                            ptrdiff_t table_sr, int npy, int npx, int C, const float *poses, const float *K,
                            int rgb_mode, int n_classes, float *tsdf, int32_t *tsdf_weight, int32_t *weight,
                            float *rgb_state, float *clip_feat, int32_t *labels, uint8_t *valid_out,
                            uint8_t *tsdf_valid_out, int64_t *counts, int num_threads, int table_mode);

int saf_oracle_integrate(const float *origin, float voxel_size, const int *nvox, int x_begin, int x_end,
                         float trunc, int B, int H, int W, const float *depth, const float *rgb,
                         const float *seg, const float *table, ptrdiff_t table_sb, ptrdiff_t table_sc,
                         ptrdiff_t table_sr, int npy, int npx, int C, const float *poses, const float *K,
                         int rgb_mode, int n_classes, float *tsdf, int32_t *tsdf_weight, int32_t *weight,
                         float *rgb_state, float *clip_feat, int32_t *labels, uint8_t *valid_out,
                         uint8_t *tsdf_valid_out, int64_t *counts, int num_threads)
{
    return saf_oracle_integrate_ex(origin, voxel_size, nvox, x_begin, x_end, trunc, B, H, W, depth, rgb, seg, table,
                                   table_sb, table_sc, table_sr, npy, npx, C, poses, K, rgb_mode, n_classes, tsdf,
                                   tsdf_weight, weight, rgb_state, clip_feat, labels, valid_out, tsdf_valid_out, counts,
                                   num_threads, 0);
}

/* table_mode 0: the reference's feature source, a [C,npy,npx] tiled-patch feature image sampled bilinearly
 * (clip_seem_fusion.py:800-805).  table_mode 1: BASELINE.json north_star's "segment -> CLIP table" (SURVEY.md 7.2,
 * NOT in the reference, so nothing pins it but this definition): table = [n_segments = npx, C] rows (npy = 1), the
 * sample of a voxel is the row of its nearest-sampled class id (the id that also feeds the label histogram,
 * clip_seem_fusion.py:786-791; 0 when the pixel is outside the image), expressed as the same four-tap chain with
 * taps (id, -, -, -) and weights (1, 0, 0, 0).  An id outside [0, n_segments) samples zeros and raises the
 * bad-label flag. */
int saf_oracle_integrate_ex(const float *origin, float voxel_size, const int *nvox, int x_begin, int x_end,
                            float trunc, int B, int H, int W, const float *depth, const float *rgb,
                            const float *seg, const float *table, ptrdiff_t table_sb, ptrdiff_t table_sc,
                            ptrdiff_t table_sr, int npy, int npx, int C, const float *poses, const float *K,
                            int rgb_mode, int n_classes, float *tsdf, int32_t *tsdf_weight, int32_t *weight,
                            float *rgb_state, float *clip_feat, int32_t *labels, uint8_t *valid_out,
                            uint8_t *tsdf_valid_out, int64_t *counts, int num_threads, int table_mode)
{
    if (table_mode == 1 && (!seg || npy != 1))
        return -3;
    if (B < 1 || B > SAF_ORACLE_MAX_BATCH)
        return -1;
    if (x_begin < 0 || x_end > nvox[0] || x_begin > x_end)
        return -2;
    const int ny = nvox[1], nz = nvox[2];
    const int64_t nslab = (int64_t)(x_end - x_begin) * ny * nz;
    const int64_t npix = (int64_t)H * W;
    int bad_label = 0;
    int64_t cnt_valid[SAF_ORACLE_MAX_BATCH] = {0}, cnt_tv[SAF_ORACLE_MAX_BATCH] = {0};
#ifdef _OPENMP
    if (num_threads > 0)
        omp_set_num_threads(num_threads);
#else
    (void)num_threads;
#endif

#pragma omp parallel for schedule(dynamic, 2048) reduction(| : bad_label) reduction(+ : cnt_valid[:SAF_ORACLE_MAX_BATCH], cnt_tv[:SAF_ORACLE_MAX_BATCH])
    for (int64_t v = 0; v < nslab; ++v) {
        const int iz = (int)(v % nz);
        const int iy = (int)((v / nz) % ny);
        const int ix = (int)(v / ((int64_t)nz * ny)) + x_begin;
        /* xyz_idx * voxel_size + origin : int64 -> f32, mul, add (clip_seem_fusion.py:669) */
        const float xw = (float)ix * voxel_size + origin[0];
        const float yw = (float)iy * voxel_size + origin[1];
        const float zw = (float)iz * voxel_size + origin[2];

        float gxs[SAF_ORACLE_MAX_BATCH], gys[SAF_ORACLE_MAX_BATCH];
        int is_valid[SAF_ORACLE_MAX_BATCH];
        int bw = 0;
        float bt = 0.0f;
        for (int b = 0; b < B; ++b) {
            float gx, gy, z;
            project(poses + 16 * b, K + 9 * b, xw, yw, zw, W, H, &gx, &gy, &z);
            int px = nearest_index(gx, W), py = nearest_index(gy, H);
            float d = (px >= 0 && py >= 0) ? depth[b * npix + (int64_t)py * W + px] : 0.0f;
            float sdf = (d - z) / trunc;
            float t = sdf < -1.0f ? -1.0f : (sdf > 1.0f ? 1.0f : sdf); /* NaN passes through */
            int in_view = (fabsf(gx) <= 1.0f) && (fabsf(gy) <= 1.0f) && (z > 0.0f);
            int valid = in_view && (fabsf(sdf) <= 1.0f);
            int tv = in_view && (sdf > -1.0f);
            gxs[b] = gx;
            gys[b] = gy;
            is_valid[b] = valid;
            if (tv) {
                bw += 1;
                bt += t;
            }
            if (valid_out)
                valid_out[b * nslab + v] = (uint8_t)valid;
            if (tsdf_valid_out)
                tsdf_valid_out[b * nslab + v] = (uint8_t)tv;
            cnt_valid[b] += valid;
            cnt_tv[b] += tv;
        }
        /* clip_seem_fusion.py:736-744 */
        const int32_t tw = tsdf_weight[v];
        const int32_t nw = tw + bw;
        if (bw > 0) {
            float a = (float)nw;
            float bb = (float)tw / (float)nw;
            tsdf[v] = bt / a + tsdf[v] * bb;
        }
        tsdf_weight[v] = nw;

        /* clip_seem_fusion.py:751-822, one image at a time */
        for (int b = 0; b < B; ++b) {
            if (!is_valid[b])
                continue;
            const float gx = gxs[b], gy = gys[b];
            const int32_t w = weight[v];
            const int32_t w1 = w + 1;
            const float a = 1.0f / (float)w1;
            const float bb = (float)w * a;

            if (seg && labels) {
                int px = nearest_index(gx, W), py = nearest_index(gy, H);
                float lf = (px >= 0 && py >= 0) ? seg[b * npix + (int64_t)py * W + px] : 0.0f;
                long lab = (long)lf;
                if (lab < 0 || lab >= n_classes)
                    bad_label |= 1;
                else
                    labels[v * (int64_t)n_classes + lab] += 1;
            }

            const float *img = rgb + b * npix * 3;
            float s[3];
            if (rgb_mode == 0) {
                int px = nearest_index(gx, W), py = nearest_index(gy, H);
                for (int k = 0; k < 3; ++k)
                    s[k] = (px >= 0 && py >= 0) ? img[((int64_t)py * W + px) * 3 + k] : 0.0f;
            } else {
                bilinear_taps t;
                bilinear_setup(gx, gy, W, H, &t);
                for (int k = 0; k < 3; ++k)
                    s[k] = bilinear_eval(&t, img + k, 3);
            }
            for (int k = 0; k < 3; ++k)
                rgb_state[v * 3 + k] = s[k] * a + rgb_state[v * 3 + k] * bb;

            bilinear_taps t;
            if (table_mode == 1) {
                int px = nearest_index(gx, W), py = nearest_index(gy, H);
                float lf = (px >= 0 && py >= 0) ? seg[b * npix + (int64_t)py * W + px] : 0.0f;
                long id = (long)lf;
                if (id < 0 || id >= npx) {
                    bad_label |= 1;
                    id = -1;
                }
                t.idx[0] = (int)id;
                t.idx[1] = t.idx[2] = t.idx[3] = -1;
                t.w[0] = 1.0f;
                t.w[1] = t.w[2] = t.w[3] = 0.0f;
            } else {
                bilinear_setup(gx, gy, npx, npy, &t);
            }
            const float *tab = table + b * table_sb;
            float *f = clip_feat + v * (int64_t)C;
            for (int c = 0; c < C; ++c) {
                float fs = bilinear_eval(&t, tab + c * table_sc, table_sr);
                f[c] = fs * a + f[c] * bb;
            }
            weight[v] = w1;
        }
    }
    if (counts)
        for (int b = 0; b < B; ++b) {
            counts[2 * b] = cnt_valid[b];
            counts[2 * b + 1] = cnt_tv[b];
        }
    return bad_label;
}

/* clip_seem_fusion.py:315-325 argmax_with_check_2d_efficient: argmax over the histogram, -1 if all zero.
 * torch.argmax returns the FIRST maximal index. */
void saf_oracle_label_argmax(const int32_t *labels, int64_t n, int n_classes, int64_t *out)
{
    for (int64_t v = 0; v < n; ++v) {
        const int32_t *row = labels + v * n_classes;
        int best = 0;
        int any = 0;
        for (int c = 0; c < n_classes; ++c) {
            if (row[c] != 0)
                any = 1;
            if (row[c] > row[best])
                best = c;
        }
        out[v] = any ? best : -1;
    }
}

int saf_oracle_max_batch(void) { return SAF_ORACLE_MAX_BATCH; }

/* Sensor-format inputs.  The reference's dataset classes (clipfusion.py:185-188, 245-254, 355-362) turn what is on
 * disk into integrate()'s fp32 arguments with   rgb = float(u8) / 255   and   depth = float(u16 mm) / 1000
 * (torch: true fp32 divisions).  The CUDA kernels take the sensor formats directly and evaluate the quotients
 * without a division: q0 = x * fl(1/d); r = fma(-d, q0, x); q = fma(r, fl(1/d), q0).  This function checks that
 * sequence against the true division for EVERY possible input and returns the number of mismatches (0). */
int saf_oracle_check_sensor_conversions(void)
{
    volatile float y1000 = 1.0f / 1000.0f, y255 = 1.0f / 255.0f;
    int bad = 0;
    for (int i = 0; i < 65536; ++i) {
        const float x = (float)i, q0 = x * y1000;
        if (fmaf(fmaf(-1000.0f, q0, x), y1000, q0) != x / 1000.0f)
            ++bad;
    }
    for (int i = 0; i < 256; ++i) {
        const float x = (float)i, q0 = x * y255;
        if (fmaf(fmaf(-255.0f, q0, x), y255, q0) != x / 255.0f)
            ++bad;
    }
    return bad;
}
