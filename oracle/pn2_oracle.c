/*
 * pn2_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked into, imported by or
 * executed from the product path).
 *
 * Scalar, single-threaded C restatement of the distance / index arithmetic of
 * the reference's PointNet++ operators (/root/reference/models/pointnet2_utils.py).
 * The reference expresses these with PyTorch CPU ops (ATen + MKL, torch
 * 2.11.0+cu128 -- un-pinned by the reference, effective pin = this image); the
 * exact fp32 operation order those ops perform was established against the
 * imported reference and is restated here explicitly so
 * that it is reproducible on any host:
 *
 *   sum(v**2,-1)          == (x*x + y*y) + z*z          separately rounded
 *   matmul, K=3           == fma(az,bz, fma(ay,by, ax*bx))
 *   square_distance       == ((-2*mm) + |src|^2) + |dst|^2
 *   FPS distance          == (dx*dx + dy*dy) + dz*dz     separately rounded
 *
 * Pinning: tests/test_oracle_golden.py checks every function below against
 * fixtures produced by the unmodified reference (tests/golden/make_golden.py).
 *
 * Build: make -C oracle   (gcc -O2 -ffp-contract=off; contraction must stay off,
 * the only fused operations are the explicit fmaf calls).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline float sq_norm3(const float *p) {
    /* torch.sum(p ** 2, -1)  -- pointnet2_utils.py:38-39 */
    float xx = p[0] * p[0];
    float yy = p[1] * p[1];
    float zz = p[2] * p[2];
    float s = xx + yy;
    return s + zz;
}

static inline float expanded_sqdist(const float *src, float s_src, const float *dst, float s_dst) {
    /* pointnet2_utils.py:37-39: -2*matmul, += |src|^2, += |dst|^2 */
    float mm = src[0] * dst[0];
    mm = fmaf(src[1], dst[1], mm);
    mm = fmaf(src[2], dst[2], mm);
    float d = -2.0f * mm;
    d = d + s_src;
    d = d + s_dst;
    return d;
}

/* a1  square_distance(src[B,N,3], dst[B,M,3]) -> out[B,N,M]   (pointnet2_utils.py:19-40) */
void pn2o_square_distance(const float *src, const float *dst, int B, int N, int M, float *out) {
    for (int b = 0; b < B; ++b) {
        const float *s = src + (size_t)b * N * 3;
        const float *d = dst + (size_t)b * M * 3;
        float *dn = (float *)malloc(sizeof(float) * (size_t)M);
        for (int j = 0; j < M; ++j) dn[j] = sq_norm3(d + 3 * j);
        for (int i = 0; i < N; ++i) {
            float sn = sq_norm3(s + 3 * i);
            float *o = out + ((size_t)b * N + i) * M;
            for (int j = 0; j < M; ++j) o[j] = expanded_sqdist(s + 3 * i, sn, d + 3 * j, dn[j]);
        }
        free(dn);
    }
}

/* a3  farthest_point_sample(xyz[B,N,3], npoint), start index supplied by the
 * caller (the reference draws it with torch.randint on the CPU generator,
 * pointnet2_utils.py:75).  Loop body = pointnet2_utils.py:77-83:
 *   distance starts at float32(1e10); dist = sum((xyz-c)**2,-1);
 *   distance = dist < distance ? dist : distance; farthest = first argmax. */
void pn2o_fps(const float *xyz, int B, int N, int npoint, const int64_t *start, int64_t *out) {
    float *mind = (float *)malloc(sizeof(float) * (size_t)N);
    for (int b = 0; b < B; ++b) {
        const float *p = xyz + (size_t)b * N * 3;
        for (int j = 0; j < N; ++j) mind[j] = 1e10f;
        int64_t far = start[b];
        for (int it = 0; it < npoint; ++it) {
            out[(size_t)b * npoint + it] = far;
            float cx = p[3 * far], cy = p[3 * far + 1], cz = p[3 * far + 2];
            float best = -1.0f;
            int64_t besti = 0;
            for (int j = 0; j < N; ++j) {
                float dx = p[3 * j] - cx;
                float dy = p[3 * j + 1] - cy;
                float dz = p[3 * j + 2] - cz;
                float xx = dx * dx, yy = dy * dy, zz = dz * dz;
                float d = xx + yy;
                d = d + zz;
                if (d < mind[j]) mind[j] = d;
                if (mind[j] > best) { best = mind[j]; besti = j; }
            }
            far = besti;
        }
    }
    free(mind);
}

/* a4  query_ball_point(radius, nsample, xyz[B,N,3], new_xyz[B,S,3])
 * (pointnet2_utils.py:87-107): ascending indices of the first nsample points
 * with !(d > r2); r2 = (float)((double)radius*radius) is computed by the caller;
 * remaining slots repeat the first hit; a query with no hit yields N in every
 * slot, as the reference's sort leaves it.  cnt (optional) = hits kept. */
void pn2o_ball_query(const float *xyz, const float *new_xyz, int B, int N, int S, float r2,
                     int nsample, int64_t *out, int32_t *cnt) {
    float *pn = (float *)malloc(sizeof(float) * (size_t)N);
    for (int b = 0; b < B; ++b) {
        const float *p = xyz + (size_t)b * N * 3;
        const float *q = new_xyz + (size_t)b * S * 3;
        for (int j = 0; j < N; ++j) pn[j] = sq_norm3(p + 3 * j);
        for (int s = 0; s < S; ++s) {
            float qn = sq_norm3(q + 3 * s);
            int64_t *o = out + ((size_t)b * S + s) * nsample;
            int c = 0;
            for (int j = 0; j < N && c < nsample; ++j) {
                float d = expanded_sqdist(q + 3 * s, qn, p + 3 * j, pn[j]);
                if (!(d > r2)) o[c++] = j;
            }
            if (cnt) cnt[(size_t)b * S + s] = c;
            int64_t first = c ? o[0] : (int64_t)N;
            for (int k = c; k < nsample; ++k) o[k] = first;
        }
    }
    free(pn);
}

/* a8 (first half)  three nearest coarse points + inverse-distance weights
 * (pointnet2_utils.py:296-302).  Stable ascending order on the expanded
 * distance (ties keep the lower index first); K3 = min(3,S) neighbours.
 *   recip = 1/(d+1e-8); norm = (r0+r1)+r2; w = recip/norm            */
void pn2o_three_nn(const float *xyz1, const float *xyz2, int B, int N, int S,
                   int64_t *idx3, float *dist3, float *w3) {
    int K3 = S < 3 ? S : 3;
    float *cn = (float *)malloc(sizeof(float) * (size_t)S);
    for (int b = 0; b < B; ++b) {
        const float *f = xyz1 + (size_t)b * N * 3;
        const float *c = xyz2 + (size_t)b * S * 3;
        for (int j = 0; j < S; ++j) cn[j] = sq_norm3(c + 3 * j);
        for (int i = 0; i < N; ++i) {
            float fn = sq_norm3(f + 3 * i);
            float bd[3] = {INFINITY, INFINITY, INFINITY};
            int64_t bi[3] = {0, 0, 0};
            int have = 0;
            for (int j = 0; j < S; ++j) {
                float d = expanded_sqdist(f + 3 * i, fn, c + 3 * j, cn[j]);
                int pos = have < K3 ? have : K3;
                while (pos > 0 && d < bd[pos - 1]) --pos;
                if (pos < K3) {
                    for (int t = K3 - 1; t > pos; --t) { bd[t] = bd[t - 1]; bi[t] = bi[t - 1]; }
                    bd[pos] = d; bi[pos] = j;
                    if (have < K3) ++have;
                }
            }
            float r[3], norm = 0.0f;
            for (int k = 0; k < K3; ++k) {
                float den = bd[k] + 1e-8f;
                r[k] = 1.0f / den;
                norm = (k == 0) ? r[0] : norm + r[k];
            }
            size_t o = ((size_t)b * N + i) * K3;
            for (int k = 0; k < K3; ++k) {
                idx3[o + k] = bi[k];
                if (dist3) dist3[o + k] = bd[k];
                w3[o + k] = r[k] / norm;
            }
        }
    }
    free(cn);
}

/* a8 (second half)  interpolated[b,n,:] = sum_k points2[b,idx_k,:] * w_k
 * (pointnet2_utils.py:303): products rounded, then added in k order. */
void pn2o_interpolate(const float *points2, const int64_t *idx3, const float *w3, int B, int N,
                      int S, int D, int K3, float *out) {
    for (int b = 0; b < B; ++b)
        for (int i = 0; i < N; ++i) {
            size_t o = ((size_t)b * N + i) * K3;
            float *dst = out + ((size_t)b * N + i) * D;
            for (int c = 0; c < D; ++c) {
                float acc = 0.0f;
                for (int k = 0; k < K3; ++k) {
                    float t = points2[((size_t)b * S + idx3[o + k]) * D + c] * w3[o + k];
                    acc = (k == 0) ? t : acc + t;
                }
                dst[c] = acc;
            }
        }
}

int pn2o_version(void) { return 1; }
