/* tb_oracle.c — CPU ORACLE (test infrastructure, see tb_oracle.h header comment).
 *
 * Plain-C restatement of the reference feature-tracking front end.  Written for clarity and
 * exactness, not speed.  Build: gcc -std=c99 -O2 -ffp-contract=off -fPIC -shared (see Makefile);
 * -ffp-contract=off because the reference is built without optimisation flags
 * (CMakeLists.txt:4-6 commented out) so no FMA contraction happens in its float math.
 */
#include "tb_oracle.h"
#include <float.h>
#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef uint8_t u8;

/* cvRound: round-half-to-even (SSE cvtsd2si / cvtss2si under the default rounding mode). */
static inline int cv_round_d(double v) { return (int)lrint(v); }
static inline int cv_round_f(float v) { return (int)lrintf(v); }
static inline int cv_floor_f(float v) { int i = (int)v; return i - (v < (float)i); }
static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }

/* ------------------------------------------------------------------------------------------
 * Frame scale tables — src/types/Frame.cpp:18-29:  sf[i] = sf[i-1] * scale  (float32 chain)
 * ------------------------------------------------------------------------------------------ */
void tbo_scale_factors(int n_levels, float scale, float *sf)
{
    for (int i = 0; i < n_levels; i++) sf[i] = 1.f;
    for (int i = 1; i < n_levels; i++) sf[i] = sf[i - 1] * scale;
}

/* Level size — src/types/Frame.cpp:424: cv::Size(image.cols * scale, image.rows * scale):
 * int * float -> float, truncated to int. */
void tbo_level_size(int w0, int h0, float sf_level, int *w, int *h)
{
    *w = (int)((float)w0 * sf_level);
    *h = (int)((float)h0 * sf_level);
}

/* ------------------------------------------------------------------------------------------
 * cv::resize(src, dst, dsize) default INTER_LINEAR, 8UC1 — call site src/types/Frame.cpp:423.
 * OpenCV (imgproc/resize.cpp) fixed-point path: 11-bit coefficients (INTER_RESIZE_COEF_SCALE =
 * 2048), horizontal pass into int32, vertical pass ((b0*(S0>>4))>>16 + (b1*(S1>>4))>>16 + 2)>>2.
 * When both scale factors are exactly 2 OpenCV silently switches to INTER_AREA (2x2 mean).
 * ------------------------------------------------------------------------------------------ */
void tbo_resize_linear_u8(const u8 *src, int sw, int sh, int sstride, u8 *dst, int dw, int dh, int dstride)
{
    double inv_scale_x = (double)dw / sw, inv_scale_y = (double)dh / sh;
    double scale_x = 1. / inv_scale_x, scale_y = 1. / inv_scale_y;
    int iscale_x = (int)lrint(scale_x), iscale_y = (int)lrint(scale_y);
    int is_area_fast = fabs(scale_x - iscale_x) < DBL_EPSILON && fabs(scale_y - iscale_y) < DBL_EPSILON;
    if (is_area_fast && iscale_x == 2 && iscale_y == 2) {
        for (int y = 0; y < dh; y++)
            for (int x = 0; x < dw; x++) {
                const u8 *s = src + (size_t)(2 * y) * sstride + 2 * x;
                dst[(size_t)y * dstride + x] = (u8)((s[0] + s[1] + s[sstride] + s[sstride + 1] + 2) >> 2);
            }
        return;
    }
    int *xofs = (int *)malloc(sizeof(int) * dw);
    short *ialpha = (short *)malloc(sizeof(short) * 2 * dw);
    int *rows = (int *)malloc(sizeof(int) * 2 * dw);
    for (int dx = 0; dx < dw; dx++) {
        float fx = (float)((dx + 0.5) * scale_x - 0.5);
        int sx = cv_floor_f(fx);
        fx -= sx;
        if (sx < 0) { fx = 0; sx = 0; }
        if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
        xofs[dx] = sx;
        ialpha[2 * dx] = (short)cv_round_f((1.f - fx) * 2048.f);
        ialpha[2 * dx + 1] = (short)cv_round_f(fx * 2048.f);
    }
    for (int dy = 0; dy < dh; dy++) {
        float fy = (float)((dy + 0.5) * scale_y - 0.5);
        int sy = cv_floor_f(fy);
        fy -= sy;
        short b0 = (short)cv_round_f((1.f - fy) * 2048.f), b1 = (short)cv_round_f(fy * 2048.f);
        int sy0 = imin(imax(sy, 0), sh - 1), sy1 = imin(imax(sy + 1, 0), sh - 1);
        const u8 *r0 = src + (size_t)sy0 * sstride, *r1 = src + (size_t)sy1 * sstride;
        for (int dx = 0; dx < dw; dx++) {
            int sx = xofs[dx], sx1 = imin(sx + 1, sw - 1);
            rows[dx] = r0[sx] * ialpha[2 * dx] + r0[sx1] * ialpha[2 * dx + 1];
            rows[dw + dx] = r1[sx] * ialpha[2 * dx] + r1[sx1] * ialpha[2 * dx + 1];
        }
        for (int dx = 0; dx < dw; dx++) {
            int v = (((b0 * (rows[dx] >> 4)) >> 16) + ((b1 * (rows[dw + dx] >> 4)) >> 16) + 2) >> 2;
            dst[(size_t)dy * dstride + dx] = (u8)v;
        }
    }
    free(xofs); free(ialpha); free(rows);
}

/* ------------------------------------------------------------------------------------------
 * cv::FAST(img, kps, T, nonmax) TYPE_9_16 — call sites src/extractors/ORBextractor.cpp:785,790.
 * OpenCV features2d/fast.cpp + fast_score.cpp: a pixel is a corner iff 9 contiguous ring pixels
 * are all > v+T or all < v-T; score = cornerScore<16> = max(A,B)-1 where A = max over arcs of the
 * min of d_k = v - ring_k, B the same for -d_k; 3-px border never tested; non-max suppression
 * keeps strict maxima of the score buffer (non-corners hold 0) over the 8-neighbourhood;
 * output row-major; KeyPoint(x, y, 7.f, -1, score).
 * ------------------------------------------------------------------------------------------ */
static const int RING_DX[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
static const int RING_DY[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

/* m = max(A,B): corner at T iff m > T; score = m - 1 (threshold independent). */
static int fast9_m(const u8 *p, int stride)
{
    int d[25];
    int v = p[0];
    for (int k = 0; k < 16; k++) d[k] = v - p[RING_DY[k] * stride + RING_DX[k]];
    for (int k = 16; k < 25; k++) d[k] = d[k - 16];
    int A = INT_MIN, B = INT_MIN;
    for (int s = 0; s < 16; s++) {
        int mn = INT_MAX, mx = INT_MIN;
        for (int k = s; k < s + 9; k++) { mn = imin(mn, d[k]); mx = imax(mx, d[k]); }
        A = imax(A, mn);
        B = imax(B, -mx);
    }
    return imax(A, B);
}

int tbo_fast9(const u8 *img, int stride, int w, int h, int T, int nonmax, tbo_kp *out, int cap)
{
    if (w < 7 || h < 7) return 0;
    int *score = (int *)calloc((size_t)w * h, sizeof(int)); /* 0 for non-corners, as OpenCV's row buffers */
    u8 *corner = (u8 *)calloc((size_t)w * h, 1);
    for (int y = 3; y < h - 3; y++)
        for (int x = 3; x < w - 3; x++) {
            const u8 *p = img + (size_t)y * stride + x;
            /* exact early reject (as OpenCV's scalar path does with its threshold table): every
             * 9-arc contains one pixel of each antipodal pair, so a corner needs one pixel of
             * (0,8) and one of (4,12) outside [v-T, v+T]. */
            int v = p[0], lo = v - T, hi = v + T;
            int a0 = p[3 * stride], a8 = p[-3 * stride];
            if (a0 >= lo && a0 <= hi && a8 >= lo && a8 <= hi) continue;
            int a4 = p[3], a12 = p[-3];
            if (a4 >= lo && a4 <= hi && a12 >= lo && a12 <= hi) continue;
            int m = fast9_m(p, stride);
            if (m > T) { score[(size_t)y * w + x] = m - 1; corner[(size_t)y * w + x] = 1; }
        }
    int n = 0;
    for (int y = 3; y < h - 3; y++)
        for (int x = 3; x < w - 3; x++) {
            if (!corner[(size_t)y * w + x]) continue;
            int s = score[(size_t)y * w + x];
            int keep = 1;
            if (nonmax) {
                for (int dy = -1; dy <= 1 && keep; dy++)
                    for (int dx = -1; dx <= 1; dx++) {
                        if (!dx && !dy) continue;
                        if (!(s > score[(size_t)(y + dy) * w + x + dx])) { keep = 0; break; }
                    }
            }
            if (!keep) continue;
            if (n < cap) {
                tbo_kp k = {(float)x, (float)y, 7.f, -1.f, (float)s, 0, -1};
                out[n] = k;
            }
            n++;
        }
    free(score); free(corner);
    return n;
}

/* ------------------------------------------------------------------------------------------
 * cv::GaussianBlur(img, img, Size(7,7), 2, 2, BORDER_REFLECT_101) on CV_8U — call sites
 * src/extractors/ORBextractor.cpp:884,960.  OpenCV >= 3.4.2 uses the bit-exact fixed-point path
 * (imgproc/smooth.dispatch.cpp, ufixedpoint16): kernel {18,34,48,56,48,34,18}/256, horizontal
 * pass in 8.8, vertical pass in 16.16, one final rounding (v + 32768) >> 16.
 * ------------------------------------------------------------------------------------------ */
static inline int reflect101(int p, int len)
{
    if (len == 1) return 0;
    while (p < 0 || p >= len) {
        if (p < 0) p = -p;
        else p = 2 * (len - 1) - p;
    }
    return p;
}

void tbo_gauss7x7(const u8 *src, int w, int h, int sstride, u8 *dst, int dstride)
{
    static const int K[7] = {18, 34, 48, 56, 48, 34, 18};
    uint16_t *tmp = (uint16_t *)malloc(sizeof(uint16_t) * (size_t)w * h);
    for (int y = 0; y < h; y++) {
        const u8 *r = src + (size_t)y * sstride;
        for (int x = 0; x < w; x++) {
            int s = 0;
            if (x >= 3 && x + 3 < w) for (int i = 0; i < 7; i++) s += K[i] * r[x + i - 3];
            else for (int i = 0; i < 7; i++) s += K[i] * r[reflect101(x + i - 3, w)];
            tmp[(size_t)y * w + x] = (uint16_t)s;
        }
    }
    for (int y = 0; y < h; y++) {
        const uint16_t *rows[7];
        for (int j = 0; j < 7; j++) rows[j] = tmp + (size_t)reflect101(y + j - 3, h) * w;
        for (int x = 0; x < w; x++) {
            uint32_t s = 0;
            for (int j = 0; j < 7; j++) s += (uint32_t)K[j] * rows[j][x];
            dst[(size_t)y * dstride + x] = (u8)((s + 32768u) >> 16);
        }
    }
    free(tmp);
}

/* ------------------------------------------------------------------------------------------
 * cv::fastAtan2(y, x) — call site src/extractors/ORBextractor.cpp:43.  OpenCV core/mathfuncs_core
 * scalar path: 7th-order odd polynomial, constants are float x float products, no FMA.
 * ------------------------------------------------------------------------------------------ */
float tbo_fast_atan2(float y, float x)
{
    const float scale = (float)(180.0 / 3.14159265358979323846);
    const float p1 = 0.9997878412794807f * scale, p3 = -0.3258083974640975f * scale;
    const float p5 = 0.1555786518463281f * scale, p7 = -0.04432655554792128f * scale;
    float ax = fabsf(x), ay = fabsf(y), a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + (float)DBL_EPSILON);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + (float)DBL_EPSILON);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

/* ------------------------------------------------------------------------------------------
 * ORBExtractor — src/extractors/ORBextractor.cpp
 * ------------------------------------------------------------------------------------------ */
#define HALF_PATCH 15
#define EDGE_THRESHOLD 19

static const int8_t ORB_PATTERN[256 * 4] = {
#include "orb_pattern.inc"
};

/* umax table — ORBextractor.cpp:389-404 */
static void make_umax(int *umax)
{
    int v, v0, vmax = cv_floor_f(HALF_PATCH * sqrtf(2.f) / 2 + 1);
    int vmin = (int)ceil(HALF_PATCH * sqrt(2.f) / 2);
    const double hp2 = HALF_PATCH * HALF_PATCH;
    for (v = 0; v <= vmax; ++v) umax[v] = cv_round_d(sqrt(hp2 - v * v));
    for (v = HALF_PATCH, v0 = 0; v >= vmin; --v) {
        while (umax[v0] == umax[v0 + 1]) ++v0;
        umax[v] = v0;
        ++v0;
    }
}

/* IC_Angle — ORBextractor.cpp:17-44 (on the un-blurred level image) */
float tbo_ic_angle(const u8 *img, int stride, int cx, int cy)
{
    int umax[HALF_PATCH + 2];
    make_umax(umax);
    int m_01 = 0, m_10 = 0;
    const u8 *center = img + (size_t)cy * stride + cx;
    for (int u = -HALF_PATCH; u <= HALF_PATCH; ++u) m_10 += u * center[u];
    for (int v = 1; v <= HALF_PATCH; ++v) {
        int v_sum = 0, d = umax[v];
        for (int u = -d; u <= d; ++u) {
            int val_plus = center[u + v * stride], val_minus = center[u - v * stride];
            v_sum += (val_plus - val_minus);
            m_10 += u * (val_plus + val_minus);
        }
        m_01 += v * v_sum;
    }
    return tbo_fast_atan2((float)m_01, (float)m_10);
}

/* computeOrbDescriptor — ORBextractor.cpp:46-87 (on the blurred level image).
 * cos/sin resolve to the float overloads (using namespace std) = glibc cosf/sinf. */
void tbo_rbrief(const u8 *blur, int stride, int cx, int cy, float angle_deg, u8 *desc)
{
    const float factorPI = (float)(3.1415926535897932384626433832795 / 180.f);
    float angle = angle_deg * factorPI;
    float a = cosf(angle), b = sinf(angle);
    const u8 *center = blur + (size_t)cy * stride + cx;
    const int8_t *pat = ORB_PATTERN;
    for (int i = 0; i < 32; ++i, pat += 32) {
        int val = 0;
        for (int k = 0; k < 8; k++) {
            const int8_t *q = pat + 4 * k;
            int t0 = center[cv_round_f(q[0] * b + q[1] * a) * stride + cv_round_f(q[0] * a - q[1] * b)];
            int t1 = center[cv_round_f(q[2] * b + q[3] * a) * stride + cv_round_f(q[2] * a - q[3] * b)];
            val |= (t0 < t1) << k;
        }
        desc[i] = (u8)val;
    }
}

/* Per-level feature budget — ORBextractor.cpp:919-930 */
void tbo_orb_budget(int n_levels, float factor, int target, int *per_level)
{
    float nDesired = target * (1 - factor) / (1 - (float)pow((double)factor, (double)n_levels));
    int sum = 0;
    for (int level = 0; level < n_levels - 1; level++) {
        per_level[level] = cv_round_f(nDesired);
        sum += per_level[level];
        nDesired *= factor;
    }
    per_level[n_levels - 1] = imax(target - sum, 0);
}

/* Cell loop of ComputeKeyPointsOctTree — ORBextractor.cpp:743-804.  Output coordinates are
 * relative to (minBorderX, minBorderY) = (16,16) exactly as vToDistributeKeys holds them. */
int tbo_cell_candidates(const u8 *img, int w, int h, int stride, float initTh, float minTh, tbo_kp *out, int cap)
{
    const float W = 30;
    const int minBorderX = EDGE_THRESHOLD - 3, minBorderY = minBorderX;
    const int maxBorderX = w - EDGE_THRESHOLD + 3, maxBorderY = h - EDGE_THRESHOLD + 3;
    const float width = (float)(maxBorderX - minBorderX), height = (float)(maxBorderY - minBorderY);
    const int nCols = (int)(width / W), nRows = (int)(height / W);
    if (nCols <= 0 || nRows <= 0) return 0;
    const int wCell = (int)ceilf(width / (float)nCols), hCell = (int)ceilf(height / (float)nRows);
    int n = 0;
    int cell_cap = (wCell + 6) * (hCell + 6);
    tbo_kp *cell = (tbo_kp *)malloc(sizeof(tbo_kp) * (size_t)cell_cap);
    for (int i = 0; i < nRows; i++) {
        const float iniY = (float)minBorderY + (float)i * (float)hCell;
        float maxY = iniY + (float)hCell + 6.f;
        if (iniY >= (float)maxBorderY - 3.f) continue;
        if (maxY > (float)maxBorderY) maxY = (float)maxBorderY;
        for (int j = 0; j < nCols; j++) {
            const float iniX = minBorderX + (float)(j * wCell);
            float maxX = iniX + (float)wCell + 6.f;
            if (iniX >= (float)maxBorderX - 6.f) continue;
            if (maxX > (float)maxBorderX) maxX = (float)maxBorderX;
            int x0 = (int)iniX, x1 = (int)maxX, y0 = (int)iniY, y1 = (int)maxY;
            const u8 *roi = img + (size_t)y0 * stride + x0;
            int nc = tbo_fast9(roi, stride, x1 - x0, y1 - y0, (int)initTh, 1, cell, cell_cap);
            if (nc == 0) nc = tbo_fast9(roi, stride, x1 - x0, y1 - y0, (int)minTh, 1, cell, cell_cap);
            for (int k = 0; k < nc; k++) {
                cell[k].x += (float)j * (float)wCell;
                cell[k].y += (float)i * (float)hCell;
                if (n < cap) out[n] = cell[k];
                n++;
            }
        }
    }
    free(cell);
    return n;
}

/* ---- DistributeOctTree — ORBextractor.cpp:494-733, ExtractorNode::DivideNode :416-491 ---- */
typedef struct {
    int ulx, urx, uly, bry; /* UL.x, UR.x, UL.y, BR.y (BL/BR.x are redundant) */
    int *keys, nkeys;       /* indices into cand, in vKeys order */
    int *ekeys, nekeys;     /* indices into exit keys */
    int no_more;
    int prev, next;         /* std::list links (-1 = end) */
} onode;

typedef struct {
    onode *n;
    int count, cap;
    int head, tail, size;
} olist;

static int ol_new(olist *L)
{
    if (L->count == L->cap) {
        L->cap = L->cap ? L->cap * 2 : 256;
        L->n = (onode *)realloc(L->n, sizeof(onode) * (size_t)L->cap);
    }
    memset(&L->n[L->count], 0, sizeof(onode));
    L->n[L->count].prev = L->n[L->count].next = -1;
    return L->count++;
}
static void ol_push_front(olist *L, int id)
{
    L->n[id].prev = -1;
    L->n[id].next = L->head;
    if (L->head >= 0) L->n[L->head].prev = id; else L->tail = id;
    L->head = id;
    L->size++;
}
static void ol_push_back(olist *L, int id)
{
    L->n[id].next = -1;
    L->n[id].prev = L->tail;
    if (L->tail >= 0) L->n[L->tail].next = id; else L->head = id;
    L->tail = id;
    L->size++;
}
static int ol_erase(olist *L, int id) /* returns next */
{
    int p = L->n[id].prev, nx = L->n[id].next;
    if (p >= 0) L->n[p].next = nx; else L->head = nx;
    if (nx >= 0) L->n[nx].prev = p; else L->tail = p;
    L->size--;
    free(L->n[id].keys); free(L->n[id].ekeys);
    L->n[id].keys = L->n[id].ekeys = NULL;
    return nx;
}

/* DivideNode: children ids written to c[4] (n1..n4). */
static void divide_node(olist *L, int id, const tbo_kp *cand, const tbo_kp *ex, int c[4])
{
    for (int k = 0; k < 4; k++) c[k] = ol_new(L);
    onode *P = &L->n[id]; /* taken after the reallocs above */
    const int halfX = (int)ceilf((float)(P->urx - P->ulx) / 2);
    const int halfY = (int)ceilf((float)(P->bry - P->uly) / 2);
    onode *n1 = &L->n[c[0]], *n2 = &L->n[c[1]], *n3 = &L->n[c[2]], *n4 = &L->n[c[3]];
    n1->ulx = P->ulx; n1->urx = P->ulx + halfX; n1->uly = P->uly; n1->bry = P->uly + halfY;
    n2->ulx = n1->urx; n2->urx = P->urx; n2->uly = P->uly; n2->bry = P->uly + halfY;
    n3->ulx = P->ulx; n3->urx = n1->urx; n3->uly = n1->bry; n3->bry = P->bry;
    n4->ulx = n1->urx; n4->urx = P->urx; n4->uly = n1->bry; n4->bry = P->bry;
    for (int k = 0; k < 4; k++) {
        L->n[c[k]].keys = (int *)malloc(sizeof(int) * (size_t)(P->nkeys + 1));
        L->n[c[k]].ekeys = (int *)malloc(sizeof(int) * (size_t)(P->nekeys + 1));
    }
    const float mx = (float)n1->urx, my = (float)n1->bry;
    for (int k = 0; k < P->nkeys; k++) {
        const tbo_kp *kp = &cand[P->keys[k]];
        onode *d = kp->x < mx ? (kp->y < my ? n1 : n3) : (kp->y < my ? n2 : n4);
        d->keys[d->nkeys++] = P->keys[k];
    }
    for (int k = 0; k < P->nekeys; k++) {
        const tbo_kp *kp = &ex[P->ekeys[k]];
        onode *d = kp->x < mx ? (kp->y < my ? n1 : n3) : (kp->y < my ? n2 : n4);
        d->ekeys[d->nekeys++] = P->ekeys[k];
    }
    for (int k = 0; k < 4; k++)
        if (L->n[c[k]].nkeys == 1) L->n[c[k]].no_more = 1;
}

typedef struct { int size, id, ulx, uly; } sz_node;
/* PINNED comparator: the reference sorts pair<int, ExtractorNode*> (ORBextractor.cpp:643), i.e.
 * ties on size are broken by heap address — not reproducible run to run.  The oracle (and
 * oracle/_ref) pin ties by (UL.x, UL.y) ascending; consumed from the back as the reference does. */
static int sz_cmp(const void *a, const void *b)
{
    const sz_node *p = (const sz_node *)a, *q = (const sz_node *)b;
    if (p->size != q->size) return p->size < q->size ? -1 : 1;
    if (p->ulx != q->ulx) return p->ulx < q->ulx ? -1 : 1;
    if (p->uly != q->uly) return p->uly < q->uly ? -1 : 1;
    return p->id < q->id ? -1 : (p->id > q->id);
}

/* split node `id`, push non-empty children to the front (n1..n4 order), collect expandable */
static void split_push(olist *L, int id, const tbo_kp *cand, const tbo_kp *ex, sz_node *exp, int *nexp)
{
    int c[4];
    divide_node(L, id, cand, ex, c);
    for (int k = 0; k < 4; k++) {
        onode *ch = &L->n[c[k]];
        if (ch->nkeys > 0) {
            ol_push_front(L, c[k]);
            if (ch->nkeys > 1) {
                sz_node s = {ch->nkeys, c[k], ch->ulx, ch->uly};
                exp[(*nexp)++] = s;
            }
        } else {
            free(ch->keys); free(ch->ekeys);
            ch->keys = ch->ekeys = NULL;
        }
    }
}

int tbo_distribute_octree(const tbo_kp *cand, int n_cand, const tbo_kp *ex, int n_exit,
                          int minX, int maxX, int minY, int maxY, int N, tbo_kp *out, int cap)
{
    const int nIni = (int)roundf((float)(maxX - minX) / (maxY - minY));
    const float hX = (float)(maxX - minX) / nIni;
    olist L = {0};
    L.head = L.tail = -1;
    if (nIni <= 0) return 0;
    for (int i = 0; i < nIni; i++) {
        int id = ol_new(&L);
        onode *ni = &L.n[id];
        ni->ulx = (int)(hX * (float)i);
        ni->urx = (int)(hX * (float)(i + 1));
        ni->uly = 0;
        ni->bry = maxY - minY;
        ni->keys = (int *)malloc(sizeof(int) * (size_t)(n_cand + 1));
        ni->ekeys = (int *)malloc(sizeof(int) * (size_t)(n_exit + 1));
        ol_push_back(&L, id);
    }
    for (int k = 0; k < n_cand; k++) {
        onode *d = &L.n[(size_t)(cand[k].x / hX)];
        d->keys[d->nkeys++] = k;
    }
    for (int k = 0; k < n_exit; k++) L.n[0].ekeys[L.n[0].nekeys++] = k;
    for (int it = L.head; it >= 0;) {
        if (L.n[it].nkeys == 1) { L.n[it].no_more = 1; it = L.n[it].next; }
        else if (L.n[it].nkeys == 0) it = ol_erase(&L, it);
        else it = L.n[it].next;
    }
    int finish = 0;
    size_t exp_cap = (size_t)(4 * (n_cand + 8));
    sz_node *exp = (sz_node *)malloc(sizeof(sz_node) * exp_cap);
    sz_node *prev = (sz_node *)malloc(sizeof(sz_node) * exp_cap);
    while (!finish) {
        int prevSize = L.size, nexp = 0;
        for (int it = L.head; it >= 0;) {
            if (L.n[it].no_more) { it = L.n[it].next; continue; }
            split_push(&L, it, cand, ex, exp, &nexp);
            it = ol_erase(&L, it);
        }
        if (L.size >= N || L.size == prevSize) finish = 1;
        else if (L.size + nexp * 3 > N) {
            while (!finish) {
                prevSize = L.size;
                int nprev = nexp;
                memcpy(prev, exp, sizeof(sz_node) * (size_t)nexp);
                nexp = 0;
                qsort(prev, (size_t)nprev, sizeof(sz_node), sz_cmp);
                for (int j = nprev - 1; j >= 0; j--) {
                    split_push(&L, prev[j].id, cand, ex, exp, &nexp);
                    ol_erase(&L, prev[j].id);
                    if (L.size >= N) break;
                }
                if (L.size >= N || L.size == prevSize) finish = 1;
            }
        }
    }
    int n = 0;
    for (int it = L.head; it >= 0; it = L.n[it].next) {
        onode *nd = &L.n[it];
        int best = nd->keys[0];
        float maxResponse = cand[best].response;
        for (int k = 1; k < nd->nkeys; k++)
            if (cand[nd->keys[k]].response > maxResponse) { best = nd->keys[k]; maxResponse = cand[best].response; }
        int accept = 1;
        for (int k = 0; k < nd->nekeys; k++) {
            const tbo_kp *e = &ex[nd->ekeys[k]];
            float dx = e->x - cand[best].x, dy = e->y - cand[best].y;
            if (dx * dx + dy * dy < 400) accept = 0;
        }
        if (accept) {
            if (n < cap) out[n] = cand[best];
            n++;
        }
    }
    for (int i = 0; i < L.count; i++) { free(L.n[i].keys); free(L.n[i].ekeys); }
    free(L.n); free(exp); free(prev);
    return n;
}

/* operator() / AddPoints — ORBextractor.cpp:906-978 / 840-904 with ComputeKeyPointsOctTree :735-828 */
int tbo_orb_extract(const u8 *const *levels, const int *w, const int *h, const int *stride, int n_levels,
                    const float *sf, int target, float initTh, float minTh,
                    const tbo_kp *exit_kps, int n_exit, const int *budget_in,
                    tbo_kp *kps, u8 *desc, int cap, int *n_per_level)
{
    if (n_levels <= 0 || !levels[0] || w[0] <= 0 || h[0] <= 0) return 0;
    int budget[64];
    if (budget_in) memcpy(budget, budget_in, sizeof(int) * (size_t)n_levels);
    else tbo_orb_budget(n_levels, sf[1], target, budget);
    int n_out = 0;
    for (int level = 0; level < n_levels; level++) {
        const int W = w[level], H = h[level], S = stride[level];
        const u8 *img = levels[level];
        const int minB = EDGE_THRESHOLD - 3, maxBX = W - EDGE_THRESHOLD + 3, maxBY = H - EDGE_THRESHOLD + 3;
        int ccap = imax((W * H) / 4, 16);
        tbo_kp *cand = (tbo_kp *)malloc(sizeof(tbo_kp) * (size_t)ccap);
        int nc = tbo_cell_candidates(img, W, H, S, initTh, minTh, cand, ccap);
        int kcap = nc + 8;
        tbo_kp *kept = (tbo_kp *)malloc(sizeof(tbo_kp) * (size_t)kcap);
        int nk = tbo_distribute_octree(cand, nc, exit_kps, n_exit, minB, maxBX, minB, maxBY, budget[level], kept, kcap);
        const int scaledPatchSize = (int)(31 * sf[level]);
        u8 *blur = NULL;
        if (nk > 0) {
            blur = (u8 *)malloc((size_t)W * H);
            tbo_gauss7x7(img, W, H, S, blur, W);
        }
        for (int i = 0; i < nk; i++) {
            tbo_kp k = kept[i];
            k.x += minB; k.y += minB;
            k.octave = level;
            k.size = (float)scaledPatchSize;
            int cx = cv_round_f(k.x), cy = cv_round_f(k.y);
            k.angle = tbo_ic_angle(img, S, cx, cy);
            if (n_out < cap) {
                tbo_rbrief(blur, W, cx, cy, k.angle, desc + (size_t)32 * n_out);
                if (level != 0) { k.x *= sf[level]; k.y *= sf[level]; } /* :969-974 (multiplies; replicated) */
                kps[n_out] = k;
            }
            n_out++;
        }
        if (n_per_level) n_per_level[level] = nk;
        free(blur); free(kept); free(cand);
    }
    return n_out;
}

/* ------------------------------------------------------------------------------------------
 * Matcher — src/matchers/matcher.cpp
 * ------------------------------------------------------------------------------------------ */
/* DescriptorDistance — matcher.cpp:793-808 (8 x 32-bit SWAR popcount) */
int tbo_hamming256(const u8 *a, const u8 *b)
{
    int dist = 0;
    for (int i = 0; i < 8; i++) {
        uint32_t pa, pb;
        memcpy(&pa, a + 4 * i, 4);
        memcpy(&pb, b + 4 * i, 4);
        uint32_t v = pa ^ pb;
        v = v - ((v >> 1) & 0x55555555u);
        v = (v & 0x33333333u) + ((v >> 2) & 0x33333333u);
        dist += (int)((((v + (v >> 4)) & 0xF0F0F0Fu) * 0x1010101u) >> 24);
    }
    return dist;
}

static inline int ham_fast(const u8 *a, const u8 *b)
{
    uint64_t x[4], y[4];
    memcpy(x, a, 32); memcpy(y, b, 32);
    return __builtin_popcountll(x[0] ^ y[0]) + __builtin_popcountll(x[1] ^ y[1]) +
           __builtin_popcountll(x[2] ^ y[2]) + __builtin_popcountll(x[3] ^ y[3]);
}

void tbo_knn2(const u8 *q, int nq, const u8 *t, int nt, int32_t *idx, int32_t *best, int32_t *second)
{
    for (int i = 0; i < nq; i++) {
        int b = INT_MAX, s = INT_MAX, bi = -1;
        for (int j = 0; j < nt; j++) {
            int d = ham_fast(q + 32 * (size_t)i, t + 32 * (size_t)j);
            if (d < b) { s = b; b = d; bi = j; }
            else if (d < s) s = d;
        }
        idx[i] = bi; best[i] = b; second[i] = s;
    }
}

/* cv::BFMatcher(NORM_HAMMING, crossCheck=true).match(d1,d2) — matcher.cpp:20,207.  OpenCV
 * features2d/matchers.cpp: knnMatch k=1 both ways (batchDistance keeps the first minimum = lowest
 * index), emit (i, nn12[i]) for ascending i iff nn21[nn12[i]] == i. */
int tbo_bf_cross(const u8 *d1, int n1, const u8 *d2, int n2, tbo_match *out)
{
    if (n1 <= 0 || n2 <= 0) return 0;
    int *nn12 = (int *)malloc(sizeof(int) * (size_t)n1), *d12 = (int *)malloc(sizeof(int) * (size_t)n1);
    int *nn21 = (int *)malloc(sizeof(int) * (size_t)n2), *d21 = (int *)malloc(sizeof(int) * (size_t)n2);
    for (int i = 0; i < n1; i++) d12[i] = INT_MAX;
    for (int j = 0; j < n2; j++) d21[j] = INT_MAX;
    for (int i = 0; i < n1; i++)
        for (int j = 0; j < n2; j++) {
            int d = ham_fast(d1 + 32 * (size_t)i, d2 + 32 * (size_t)j);
            if (d < d12[i]) { d12[i] = d; nn12[i] = j; }
            if (d < d21[j]) { d21[j] = d; nn21[j] = i; }
        }
    int n = 0;
    for (int i = 0; i < n1; i++)
        if (nn21[nn12[i]] == i) {
            tbo_match m = {i, nn12[i], 0, (float)d12[i]};
            out[n++] = m;
        }
    free(nn12); free(d12); free(nn21); free(d21);
    return n;
}

/* searchByBF, all-levels branch — matcher.cpp:168-228 */
int tbo_search_by_bf(const u8 *d1, int n1, const u8 *d2, int n2, float ratio, float minTh, tbo_match *out)
{
    tbo_match *all = (tbo_match *)malloc(sizeof(tbo_match) * (size_t)(n1 > 0 ? n1 : 1));
    int na = tbo_bf_cross(d1, n1, d2, n2, all);
    int n = 0;
    if (na > 0) { /* the reference dereferences min_element of an empty vector otherwise (UB) */
        float min_distance = all[0].distance;
        for (int i = 1; i < na; i++) if (all[i].distance < min_distance) min_distance = all[i].distance;
        for (int i = 0; i < na; i++)
            if (all[i].distance < fminf(ratio * min_distance, minTh)) out[n++] = all[i];
    }
    free(all);
    return n;
}

/* ---- Frame grid — src/types/Frame.cpp:30-31 (scale members), 187-200, 202-255, 257-265 ---- */
#define GRID_ROWS 36
#define GRID_COLS 120
struct tbo_grid {
    float wInv, hInv; /* mfGridElementWidthInv = 36/rows, mfGridElementHeightInv = 120/cols (sic) */
    int *cell_start;  /* [GRID_COLS*GRID_ROWS + 1], cell (ix,iy) at ix*GRID_ROWS+iy */
    int *items;
    const tbo_kp *kps;
    int n;
};

static int pos_in_grid(const tbo_grid *g, const tbo_kp *kp, int *px, int *py)
{
    *px = (int)roundf(kp->x * g->wInv);
    *py = (int)roundf(kp->y * g->hInv);
    return !(*px < 0 || *px >= GRID_COLS || *py < 0 || *py >= GRID_ROWS);
}

tbo_grid *tbo_grid_build(const tbo_kp *kps, int n, int img_cols, int img_rows)
{
    tbo_grid *g = (tbo_grid *)calloc(1, sizeof(tbo_grid));
    g->hInv = (float)GRID_COLS / (float)img_cols;
    g->wInv = (float)GRID_ROWS / (float)img_rows;
    g->kps = kps; g->n = n;
    int ncell = GRID_COLS * GRID_ROWS;
    g->cell_start = (int *)calloc((size_t)ncell + 1, sizeof(int));
    g->items = (int *)malloc(sizeof(int) * (size_t)(n + 1));
    int *cnt = (int *)calloc((size_t)ncell, sizeof(int));
    for (int i = 0; i < n; i++) { int x, y; if (pos_in_grid(g, &kps[i], &x, &y)) cnt[x * GRID_ROWS + y]++; }
    for (int c = 0; c < ncell; c++) g->cell_start[c + 1] = g->cell_start[c] + cnt[c];
    memset(cnt, 0, sizeof(int) * (size_t)ncell);
    for (int i = 0; i < n; i++) {
        int x, y;
        if (pos_in_grid(g, &kps[i], &x, &y)) { int c = x * GRID_ROWS + y; g->items[g->cell_start[c] + cnt[c]++] = i; }
    }
    free(cnt);
    return g;
}
void tbo_grid_free(tbo_grid *g) { if (g) { free(g->cell_start); free(g->items); free(g); } }

int tbo_features_in_area(const tbo_grid *g, float x, float y, float r, int minLevel, int maxLevel, int32_t *out, int cap)
{
    int n = 0;
    const int nMinCellX = imax(0, (int)floorf((x - r) * g->wInv));
    if (nMinCellX >= GRID_COLS) return 0;
    const int nMaxCellX = imin(GRID_COLS - 1, (int)ceilf((x + r) * g->wInv));
    if (nMaxCellX < 0) return 0;
    const int nMinCellY = imax(0, (int)floorf((y - r) * g->hInv));
    if (nMinCellY >= GRID_ROWS) return 0;
    const int nMaxCellY = imin(GRID_ROWS - 1, (int)ceilf((y + r) * g->hInv));
    if (nMaxCellY < 0) return 0;
    const int bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
    for (int ix = nMinCellX; ix <= nMaxCellX; ix++)
        for (int iy = nMinCellY; iy <= nMaxCellY; iy++) {
            int c = ix * GRID_ROWS + iy;
            for (int k = g->cell_start[c]; k < g->cell_start[c + 1]; k++) {
                int j = g->items[k];
                const tbo_kp *kp = &g->kps[j];
                if (bCheckLevels) {
                    if (kp->octave < minLevel) continue;
                    if (maxLevel >= 0 && kp->octave > maxLevel) continue;
                }
                const float distx = kp->x - x, disty = kp->y - y;
                if (fabsf(distx) < r && fabsf(disty) < r) { if (n < cap) out[n] = j; n++; }
            }
        }
    return n;
}

/* ComputeThreeMaxima — matcher.cpp:810-851 */
void tbo_three_maxima(const int *hs, int L, int *ind1, int *ind2, int *ind3)
{
    int max1 = 0, max2 = 0, max3 = 0;
    for (int i = 0; i < L; i++) {
        const int s = hs[i];
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; *ind3 = *ind2; *ind2 = *ind1; *ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; *ind3 = *ind2; *ind2 = i; }
        else if (s > max3) { max3 = s; *ind3 = i; }
    }
    if ((float)max2 < 0.1f * (float)max1) { *ind2 = -1; *ind3 = -1; }
    else if ((float)max3 < 0.1f * (float)max1) { *ind3 = -1; }
}

/* rotation-histogram regrouping shared by searchByViolence (:359-393) and searchByProjection
 * (:493-529): bins[i] = histogram bin of raw match i; output = matches grouped by bin ascending,
 * only the three maxima bins. */
static int rot_hist_filter(const tbo_match *raw, const int *bins, int n, int L, tbo_match *out)
{
    int *hs = (int *)calloc((size_t)L, sizeof(int));
    for (int i = 0; i < n; i++) hs[bins[i]]++;
    int ind[3] = {-1, -1, -1};
    tbo_three_maxima(hs, L, &ind[0], &ind[1], &ind[2]);
    int m = 0;
    for (int b = 0; b < L; b++)
        if (b == ind[0] || b == ind[1] || b == ind[2])
            for (int i = 0; i < n; i++) if (bins[i] == b) out[m++] = raw[i];
    free(hs);
    return m;
}

/* searchByViolence — matcher.cpp:299-395 (MapPointOnly == false) */
int tbo_search_by_violence(const tbo_matcher_params *p, const tbo_kp *k1, const u8 *d1, int n1,
                           const tbo_kp *k2, const u8 *d2, int n2, int cols2, int rows2,
                           int min_level, int max_level, float search_r, tbo_match *out)
{
    tbo_grid *g = tbo_grid_build(k2, n2, cols2, rows2);
    int32_t *ind = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n2 + 1));
    tbo_match *raw = (tbo_match *)malloc(sizeof(tbo_match) * (size_t)(n1 + 1));
    int *bins = (int *)malloc(sizeof(int) * (size_t)(n1 + 1));
    const float factor = 1.f / (float)p->HISTO_LENGTH;
    int n = 0;
    for (int i1 = 0; i1 < n1; i1++) {
        int nc = tbo_features_in_area(g, k1[i1].x, k1[i1].y, search_r, min_level, max_level, ind, n2);
        if (nc == 0) continue;
        int bestDist = INT_MAX, bestDist2 = INT_MAX, bestIdx2 = -1;
        for (int c = 0; c < nc; c++) {
            int dist = tbo_hamming256(d1 + 32 * (size_t)i1, d2 + 32 * (size_t)ind[c]);
            if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestIdx2 = ind[c]; }
            else if (dist < bestDist2) bestDist2 = dist;
        }
        if (bestDist <= p->TH_LOW && (float)bestDist < (float)bestDist2 * p->nRatio) {
            tbo_match m = {i1, bestIdx2, -1, (float)bestDist};
            raw[n] = m;
            if (p->checkOrientation) {
                float rot = k1[i1].angle - k2[bestIdx2].angle;
                if (rot < 0) rot += 360.f;
                int bin = (int)roundf(rot * factor);
                if (bin == p->HISTO_LENGTH) bin = 0;
                bins[n] = bin;
            }
            n++;
        }
    }
    int m;
    if (p->checkOrientation) m = rot_hist_filter(raw, bins, n, p->HISTO_LENGTH, out);
    else { memcpy(out, raw, sizeof(tbo_match) * (size_t)n); m = n; }
    free(ind); free(raw); free(bins);
    tbo_grid_free(g);
    return m;
}

/* ---- frame ingest ------------------------------------------------------------------------------
 * cv::initUndistortRectifyMap (OpenCV imgproc/undistort.dispatch.cpp): R = I, new camera matrix = K, so
 * ir = K^-1 by cv::invert's closed-form 3x3 (cofactors * 1/det, double); per row the homogeneous (x,y,w)
 * is advanced by column 0 of ir; radial k1,k2,k3 + tangential p1,p2; result rounded to float. */
void tbo_init_undistort_rectify_map(float ffx, float ffy, float fcx, float fcy, const float d[5], int w, int h,
                                    float *map1, float *map2)
{
    const double S[3][3] = {{ffx, 0, fcx}, {0, ffy, fcy}, {0, 0, 1}};
    double det = S[0][0] * (S[1][1] * S[2][2] - S[1][2] * S[2][1]) - S[0][1] * (S[1][0] * S[2][2] - S[1][2] * S[2][0]) +
                 S[0][2] * (S[1][0] * S[2][1] - S[1][1] * S[2][0]);
    det = 1.0 / det;
    double ir[9];
    ir[0] = (S[1][1] * S[2][2] - S[1][2] * S[2][1]) * det;
    ir[1] = (S[0][2] * S[2][1] - S[0][1] * S[2][2]) * det;
    ir[2] = (S[0][1] * S[1][2] - S[0][2] * S[1][1]) * det;
    ir[3] = (S[1][2] * S[2][0] - S[1][0] * S[2][2]) * det;
    ir[4] = (S[0][0] * S[2][2] - S[0][2] * S[2][0]) * det;
    ir[5] = (S[0][2] * S[1][0] - S[0][0] * S[1][2]) * det;
    ir[6] = (S[1][0] * S[2][1] - S[1][1] * S[2][0]) * det;
    ir[7] = (S[0][1] * S[2][0] - S[0][0] * S[2][1]) * det;
    ir[8] = (S[0][0] * S[1][1] - S[0][1] * S[1][0]) * det;
    const double fx = ffx, fy = ffy, u0 = fcx, v0 = fcy;
    const double k1 = d[0], k2 = d[1], p1 = d[2], p2 = d[3], k3 = d[4];
    for (int i = 0; i < h; i++) {
        double _x = i * ir[1] + ir[2], _y = i * ir[4] + ir[5], _w = i * ir[7] + ir[8];
        for (int j = 0; j < w; j++, _x += ir[0], _y += ir[3], _w += ir[6]) {
            const double ww = 1. / _w, x = _x * ww, y = _y * ww;
            const double x2 = x * x, y2 = y * y, r2 = x2 + y2, _2xy = 2 * x * y;
            const double kr = (1 + ((k3 * r2 + k2) * r2 + k1) * r2) / (1 + ((0 * r2 + 0) * r2 + 0) * r2);
            const double xd = (x * kr + p1 * _2xy + p2 * (r2 + 2 * x2) + 0 * r2 + 0 * r2 * r2);
            const double yd = (y * kr + p1 * (r2 + 2 * y2) + p2 * _2xy + 0 * r2 + 0 * r2 * r2);
            map1[(size_t)i * w + j] = (float)(fx * 1.0 * xd + u0);
            map2[(size_t)i * w + j] = (float)(fy * 1.0 * yd + v0);
        }
    }
}

/* cvRound(float) as SSE cvtss2si computes it: NaN / out-of-range give INT_MIN ("integer indefinite") */
static inline int cv_round_f_indef(float v)
{
    if (!(v >= -2147483648.f && v < 2147483648.f)) return INT_MIN;
    return (int)lrintf(v);
}
static inline int sat_short(int v) { return v < -32768 ? -32768 : (v > 32767 ? 32767 : v); }

void tbo_remap_linear_u8(const u8 *src, int sw, int sh, int sstride, const float *map1, const float *map2,
                         int dw, int dh, u8 *dst, int dstride)
{
    for (int y = 0; y < dh; y++)
        for (int x = 0; x < dw; x++) {
            const int sx = cv_round_f_indef(map1[(size_t)y * dw + x] * 32.f), sy = cv_round_f_indef(map2[(size_t)y * dw + x] * 32.f);
            const int fx = sx & 31, fy = sy & 31;
            const int ix = sat_short(sx >> 5), iy = sat_short(sy >> 5);
            int v[4];
            for (int t = 0; t < 4; t++) {
                const int xx = ix + (t & 1), yy = iy + (t >> 1);
                v[t] = (xx >= 0 && xx < sw && yy >= 0 && yy < sh) ? src[(size_t)yy * sstride + xx] : 0;
            }
            const int sum = v[0] * ((32 - fy) * (32 - fx) * 32) + v[1] * ((32 - fy) * fx * 32) + v[2] * (fy * (32 - fx) * 32) + v[3] * (fy * fx * 32);
            const int r = (sum + (1 << 14)) >> 15;
            dst[(size_t)y * dstride + x] = (u8)(r < 0 ? 0 : (r > 255 ? 255 : r));
        }
}

/* cv::CLAHE (OpenCV imgproc/clahe.cpp), 8-bit: tile LUTs from clipped + redistributed histograms of the image
 * extended (BORDER_REFLECT_101, right/bottom) to a multiple of the tile grid, then per-pixel bilinear blend of
 * the four neighbouring tiles' LUT values in float, saturate_cast<uchar> (round-half-even). */
void tbo_clahe_u8(const u8 *src, int w, int h, int stride, double clip_limit, int tiles_x, int tiles_y, u8 *dst, int dstride)
{
    const int ew = (w % tiles_x == 0 && h % tiles_y == 0) ? w : w + (tiles_x - w % tiles_x);
    const int eh = (w % tiles_x == 0 && h % tiles_y == 0) ? h : h + (tiles_y - h % tiles_y);
    const int tw = ew / tiles_x, th = eh / tiles_y, area = tw * th;
    const float lutScale = (float)255 / (float)area;
    int clip = 0;
    if (clip_limit > 0.0) { clip = (int)(clip_limit * area / 256); clip = imax(clip, 1); }
    u8 *lut = (u8 *)malloc((size_t)tiles_x * tiles_y * 256);
    for (int ty = 0; ty < tiles_y; ty++)
        for (int tx = 0; tx < tiles_x; tx++) {
            int hist[256];
            memset(hist, 0, sizeof hist);
            for (int y = ty * th; y < (ty + 1) * th; y++)
                for (int x = tx * tw; x < (tx + 1) * tw; x++)
                    hist[src[(size_t)reflect101(y, h) * stride + reflect101(x, w)]]++;
            if (clip > 0) {
                int clipped = 0;
                for (int i = 0; i < 256; i++) if (hist[i] > clip) { clipped += hist[i] - clip; hist[i] = clip; }
                const int batch = clipped / 256;
                int residual = clipped - batch * 256;
                for (int i = 0; i < 256; i++) hist[i] += batch;
                if (residual != 0) {
                    const int step = imax(256 / residual, 1);
                    for (int i = 0; i < 256 && residual > 0; i += step, residual--) hist[i]++;
                }
            }
            u8 *l = lut + ((size_t)ty * tiles_x + tx) * 256;
            int sum = 0;
            for (int i = 0; i < 256; i++) {
                sum += hist[i];
                const int r = cv_round_f((float)sum * lutScale);
                l[i] = (u8)(r < 0 ? 0 : (r > 255 ? 255 : r));
            }
        }
    const float inv_tw = 1.0f / (float)tw, inv_th = 1.0f / (float)th;
    for (int y = 0; y < h; y++) {
        const float tyf = (float)y * inv_th - 0.5f;
        int ty1 = cv_floor_f(tyf), ty2 = ty1 + 1;
        const float ya = tyf - (float)ty1, ya1 = 1.0f - ya;
        ty1 = imax(ty1, 0); ty2 = imin(ty2, tiles_y - 1);
        for (int x = 0; x < w; x++) {
            const float txf = (float)x * inv_tw - 0.5f;
            int tx1 = cv_floor_f(txf), tx2 = tx1 + 1;
            const float xa = txf - (float)tx1, xa1 = 1.0f - xa;
            tx1 = imax(tx1, 0); tx2 = imin(tx2, tiles_x - 1);
            const int v = src[(size_t)y * stride + x];
            const float a = lut[((size_t)ty1 * tiles_x + tx1) * 256 + v], b = lut[((size_t)ty1 * tiles_x + tx2) * 256 + v];
            const float c = lut[((size_t)ty2 * tiles_x + tx1) * 256 + v], e = lut[((size_t)ty2 * tiles_x + tx2) * 256 + v];
            const float res = (a * xa1 + b * xa) * ya1 + (c * xa1 + e * xa) * ya;
            const int r = cv_round_f(res);
            dst[(size_t)y * dstride + x] = (u8)(r < 0 ? 0 : (r > 255 ? 255 : r));
        }
    }
    free(lut);
}

/* TemplatedVocabulary::transform per feature — TemplatedVocabulary.h:1218-1259; FORB::distance is the
 * 256-bit Hamming distance (FORB.cpp:81-101); first minimum wins (strict d < best_d). */
void tbo_bow_transform(const tbo_vocab *v, const u8 *desc, int n, int levelsup,
                       int32_t *word_id, int32_t *node_id, double *weight)
{
    const int nid_level = v->L - levelsup;
    for (int i = 0; i < n; i++) {
        const u8 *f = desc + 32 * (size_t)i;
        int nid = (nid_level <= 0) ? 0 : -1;
        int final_id = 0, level = 0;
        do {
            level++;
            const int cs = v->child_start[final_id], ce = v->child_start[final_id + 1];
            int best = v->children[cs];
            int best_d = tbo_hamming256(f, v->desc + 32 * (size_t)best);
            for (int c = cs + 1; c < ce; c++) {
                const int id = v->children[c];
                const int d = tbo_hamming256(f, v->desc + 32 * (size_t)id);
                if (d < best_d) { best_d = d; best = id; }
            }
            final_id = best;
            if (level == nid_level) nid = final_id;
        } while (v->child_start[final_id + 1] > v->child_start[final_id]);
        if (nid < 0) nid = final_id; /* leaf above nid_level: UB in the reference, pinned to the leaf */
        word_id[i] = v->word_id[final_id];
        node_id[i] = nid;
        weight[i] = v->weight[final_id];
    }
}

/* searchByBow — matcher.cpp:619-721.  The two FeatureVectors are std::map<NodeId, vector<idx>>: walk the common
 * node ids in ascending order, members in insertion (feature index) order. */
static int bow_cmp(const void *a, const void *b)
{
    const int64_t x = *(const int64_t *)a, y = *(const int64_t *)b;
    return (x > y) - (x < y);
}
int tbo_search_by_bow(const tbo_matcher_params *p, const tbo_kp *k1, const u8 *d1, const int32_t *node1, int n1,
                      const tbo_kp *k2, const u8 *d2, const int32_t *node2, int n2,
                      const u8 *mp_valid2, int map_point_only, tbo_match *out)
{
    int64_t *o1 = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n1 + 1));
    int64_t *o2 = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n2 + 1));
    tbo_match *raw = (tbo_match *)malloc(sizeof(tbo_match) * (size_t)(n1 + 1));
    int *bins = (int *)malloc(sizeof(int) * (size_t)(n1 + 1));
    int m1 = 0, m2 = 0, n = 0;
    for (int i = 0; i < n1; i++) if (node1[i] >= 0) o1[m1++] = ((int64_t)node1[i] << 32) | (uint32_t)i;
    for (int i = 0; i < n2; i++) if (node2[i] >= 0) o2[m2++] = ((int64_t)node2[i] << 32) | (uint32_t)i;
    qsort(o1, (size_t)m1, sizeof(int64_t), bow_cmp);
    qsort(o2, (size_t)m2, sizeof(int64_t), bow_cmp);
    const float factor = 1.f / (float)p->HISTO_LENGTH;
    int a = 0, b = 0;
    while (a < m1 && b < m2) {
        const int na = (int)(o1[a] >> 32), nb = (int)(o2[b] >> 32);
        if (na < nb) { a++; continue; }
        if (nb < na) { b++; continue; }
        int ae = a, be = b;
        while (ae < m1 && (int)(o1[ae] >> 32) == na) ae++;
        while (be < m2 && (int)(o2[be] >> 32) == na) be++;
        for (int x = a; x < ae; x++) {
            const int idx1 = (int)(o1[x] & 0xffffffff);
            int bestDist1 = 256, bestIdx2 = -1, bestDist2 = 256;
            for (int y = b; y < be; y++) {
                const int idx2 = (int)(o2[y] & 0xffffffff);
                if (map_point_only && !(mp_valid2 && mp_valid2[idx2])) continue;
                const int dist = tbo_hamming256(d1 + 32 * (size_t)idx1, d2 + 32 * (size_t)idx2);
                if (dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestIdx2 = idx2; }
                else if (dist < bestDist2) bestDist2 = dist;
            }
            if (bestIdx2 >= 0 && bestDist1 < p->TH_LOW && (float)bestDist1 < p->nRatio * (float)bestDist2) {
                tbo_match m = {idx1, bestIdx2, -1, (float)bestDist1};
                raw[n] = m;
                if (p->checkOrientation) {
                    float rot = k1[idx1].angle - k2[bestIdx2].angle;
                    if (rot < 0) rot += 360.f;
                    int bin = (int)roundf(rot * factor);
                    if (bin == p->HISTO_LENGTH) bin = 0;
                    bins[n] = bin;
                }
                n++;
            }
        }
        a = ae; b = be;
    }
    int m;
    if (p->checkOrientation) m = rot_hist_filter(raw, bins, n, p->HISTO_LENGTH, out);
    else { memcpy(out, raw, sizeof(tbo_match) * (size_t)n); m = n; }
    free(o1); free(o2); free(raw); free(bins);
    return m;
}

/* PinholeCamera::World2Cam without distortion — CameraModel.cpp:63-74; IsInFrame on truncated
 * ints — CameraModel.h:33-39 (boundary 0, scale 1). */
static void world2cam(const tbo_pinhole *c, const float *xyz, float *u, float *v)
{
    float x = xyz[0] / xyz[2], y = xyz[1] / xyz[2];
    if (!(fabs(c->d[0]) > 0.0000001)) { /* mbDistortion — CameraModel.cpp:9 */
        *u = c->fx * x + c->cx;
        *v = c->fy * y + c->cy;
    } else { /* radtan — CameraModel.cpp:75-92 */
        float r2 = x * x + y * y, r4 = r2 * r2, r6 = r4 * r2;
        float a1 = 2 * x * y, a2 = r2 + 2 * x * x, a3 = r2 + 2 * y * y;
        float cdist = 1 + c->d[0] * r2 + c->d[1] * r4 + c->d[4] * r6;
        float xd = x * cdist + c->d[2] * a1 + c->d[3] * a2;
        float yd = y * cdist + c->d[2] * a3 + c->d[3] * a1;
        *u = xd * c->fx + c->cx;
        *v = yd * c->fy + c->cy;
    }
}
static int is_in_frame(const tbo_pinhole *c, float u, float v)
{
    int iu = (int)u, iv = (int)v;
    return iu >= 0 && iu < (int)((float)c->width * 1.f) && iv >= 0 && iv < (int)((float)c->height * 1.f);
}
/* Eigen evaluates a fixed-size 3-term dot product with its non-vectorised redux unroller
 * (Eigen/src/Core/Redux.h, HalfLength = 3/2 = 1): x0 + (x1 + x2).  Eigen is an un-vendored
 * dependency of the reference; this order is restated here and mirrored by the CUDA path. */
static inline float dot3(float a0, float b0, float a1, float b1, float a2, float b2)
{
    return a0 * b0 + (a1 * b1 + a2 * b2);
}
static void transform_pt(const float *T, const float *X, float *out) /* Rcw*Xw + tcw */
{
    for (int r = 0; r < 3; r++)
        out[r] = dot3(T[4 * r + 0], X[0], T[4 * r + 1], X[1], T[4 * r + 2], X[2]) + T[4 * r + 3];
}

/* searchByProjection(F1,F2) — matcher.cpp:405-531 */
int tbo_search_by_projection(const tbo_matcher_params *p, const tbo_pinhole *cam, const float *Tcw1,
                             const tbo_kp *k1, const u8 *d1, int n1, const u8 *f1_matched,
                             const float *sf1, int n_levels1, int cols1, int rows1,
                             const tbo_kp *k2, int n2, const u8 *mp_valid, const float *mp_xyz,
                             const u8 *mp_desc, tbo_match *out)
{
    (void)n_levels1;
    tbo_grid *g = tbo_grid_build(k1, n1, cols1, rows1);
    int32_t *ind = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n1 + 1));
    tbo_match *raw = (tbo_match *)malloc(sizeof(tbo_match) * (size_t)(n2 + 1));
    int *bins = (int *)malloc(sizeof(int) * (size_t)(n2 + 1));
    const float factor = 1.0f / (float)p->HISTO_LENGTH;
    int n = 0;
    for (int i2 = 0; i2 < n2; i2++) {
        if (!mp_valid[i2]) continue;
        float xc[3];
        transform_pt(Tcw1, mp_xyz + 3 * (size_t)i2, xc);
        const float invzc = 1.0f / xc[2];
        if (invzc < 0) continue;
        float u, v;
        world2cam(cam, xc, &u, &v);
        if (!is_in_frame(cam, u, v)) continue;
        int nLastOctave = k2[i2].octave;
        float radius = p->nRatio * sf1[nLastOctave];
        int nc = tbo_features_in_area(g, u, v, radius, nLastOctave - 1, nLastOctave + 1, ind, n1);
        if (nc == 0) continue;
        int bestDist = 256, bestIdx1 = -1;
        for (int c = 0; c < nc; c++) {
            int i1 = ind[c];
            if (f1_matched && f1_matched[i1]) continue;
            int dist = tbo_hamming256(mp_desc + 32 * (size_t)i2, d1 + 32 * (size_t)i1);
            if (dist < bestDist) { bestDist = dist; bestIdx1 = i1; }
        }
        if (bestDist <= p->TH_HIGH) {
            tbo_match m = {bestIdx1, i2, -1, (float)bestDist};
            raw[n] = m;
            if (p->checkOrientation) {
                float rot = k2[i2].angle - k1[bestIdx1].angle;
                if (rot < 0.0) rot += 360.0f;
                int bin = (int)round(rot * factor);
                if (bin == p->HISTO_LENGTH) bin = 0;
                bins[n] = bin;
            }
            n++;
        }
    }
    int m;
    if (p->checkOrientation) m = rot_hist_filter(raw, bins, n, p->HISTO_LENGTH, out);
    else { memcpy(out, raw, sizeof(tbo_match) * (size_t)n); m = n; }
    free(ind); free(raw); free(bins);
    tbo_grid_free(g);
    return m;
}

/* searchByProjection(map,F1,radio) — matcher.cpp:539-617; Frame::IsInFrustum — Frame.cpp:370-412 */
int tbo_search_by_projection_map(const tbo_matcher_params *p, const tbo_pinhole *cam, const float *Tcw1,
                                 const tbo_kp *k1, const u8 *d1, int n1, const u8 *f1_matched,
                                 const float *sf1, int cols1, int rows1,
                                 int n_mp, const u8 *mp_bad, const float *mp_xyz, const float *mp_normal,
                                 const u8 *mp_desc, float radio, tbo_match *out)
{
    tbo_grid *g = tbo_grid_build(k1, n1, cols1, rows1);
    int32_t *ind = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n1 + 1));
    const int bFactor = p->nRatio != 1.0;
    /* camera centre mOw = -Rcw^T tcw — Frame.cpp:57 */
    float Ow[3];
    for (int c = 0; c < 3; c++)
        Ow[c] = dot3(-Tcw1[0 + c], Tcw1[3], -Tcw1[4 + c], Tcw1[7], -Tcw1[8 + c], Tcw1[11]);
    int n = 0, cnt = 0;
    for (int im = 0; im < n_mp; im++) {
        cnt++;
        if (mp_bad[im]) continue;
        const float *P = mp_xyz + 3 * (size_t)im;
        float Pc[3];
        transform_pt(Tcw1, P, Pc);
        if (Pc[2] < 0.0f) continue;
        float u, v;
        world2cam(cam, Pc, &u, &v);
        if (!is_in_frame(cam, u, v)) continue;
        const float maxDistance = 1000.f, minDistance = 1.f; /* MapPoint.cpp:207-217 */
        float PO[3] = {P[0] - Ow[0], P[1] - Ow[1], P[2] - Ow[2]};
        const float dist3 = sqrtf(dot3(PO[0], PO[0], PO[1], PO[1], PO[2], PO[2]));
        if (dist3 < minDistance || dist3 > maxDistance) continue;
        const float *Pn = mp_normal + 3 * (size_t)im;
        const float viewCos = dot3(PO[0], Pn[0], PO[1], Pn[1], PO[2], Pn[2]) / dist3;
        if (viewCos < 0.5f) continue;
        const int nPredictedLevel = 0;
        float r = 4.f;
        if (viewCos > 0.998) r = 2.5;
        if (bFactor) r *= p->nRatio;
        int nc = tbo_features_in_area(g, u, v, r * sf1[nPredictedLevel], nPredictedLevel - 1, nPredictedLevel, ind, n1);
        if (nc == 0) continue;
        int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
        for (int c = 0; c < nc; c++) {
            int idx = ind[c];
            if (f1_matched && f1_matched[idx]) continue;
            const int dist = tbo_hamming256(mp_desc + 32 * (size_t)im, d1 + 32 * (size_t)idx);
            if (dist < bestDist) {
                bestDist2 = bestDist; bestDist = dist; bestLevel2 = bestLevel;
                bestLevel = k1[idx].octave; bestIdx = idx;
            } else if (dist < bestDist2) { bestLevel2 = k1[idx].octave; bestDist2 = dist; }
        }
        if (bestDist <= p->TH_HIGH) {
            if (bestLevel == bestLevel2 && bestDist > radio * bestDist2) continue;
            tbo_match m = {bestIdx, cnt - 1, -1, (float)bestDist};
            out[n++] = m;
        }
    }
    free(ind);
    tbo_grid_free(g);
    return n;
}

/* FASTExtractor::shiTomasiScore — src/extractors/FASTextractor.cpp:87-127 */
float tbo_shi_tomasi(const u8 *img, int w, int h, int stride, int u, int v)
{
    float dXX = 0.0, dYY = 0.0, dXY = 0.0;
    const int halfbox_size = 4, box_size = 2 * halfbox_size, box_area = box_size * box_size;
    const int x_min = u - halfbox_size, x_max = u + halfbox_size;
    const int y_min = v - halfbox_size, y_max = v + halfbox_size;
    if (x_min < 1 || x_max >= w - 1 || y_min < 1 || y_max >= h - 1) return 0.0;
    for (int y = y_min; y < y_max; ++y) {
        const u8 *pl = img + (size_t)stride * y + x_min - 1, *pr = img + (size_t)stride * y + x_min + 1;
        const u8 *pt = img + (size_t)stride * (y - 1) + x_min, *pb = img + (size_t)stride * (y + 1) + x_min;
        for (int x = 0; x < box_size; ++x, ++pl, ++pr, ++pt, ++pb) {
            float dx = (float)*pr - (float)*pl, dy = (float)*pb - (float)*pt;
            dXX += dx * dx; dYY += dy * dy; dXY += dx * dy;
        }
    }
    dXX = dXX / (2.f * box_area); dYY = dYY / (2.f * box_area); dXY = dXY / (2.f * box_area);
    return 0.5f * (dXX + dYY - sqrtf((dXX + dYY) * (dXX + dYY) - 4 * (dXX * dYY - dXY * dXY)));
}

/* ------------------------------------------------------------------------------------------
 * FASTExtractor — src/extractors/FASTextractor.cpp (SVO-style grid FAST).  Detector restated from the
 * public uzh-rpg/fast (Rosten) algorithm; see tb_oracle.h: parity unpinned.
 * ------------------------------------------------------------------------------------------ */
static int fast10_m(const u8 *p, int stride)
{
    int d[26];
    int v = p[0];
    for (int k = 0; k < 16; k++) d[k] = v - p[RING_DY[k] * stride + RING_DX[k]];
    for (int k = 16; k < 26; k++) d[k] = d[k - 16];
    int A = INT_MIN, B = INT_MIN;
    for (int s = 0; s < 16; s++) {
        int mn = INT_MAX, mx = INT_MIN;
        for (int k = s; k < s + 10; k++) { mn = imin(mn, d[k]); mx = imax(mx, d[k]); }
        A = imax(A, mn);
        B = imax(B, -mx);
    }
    return imax(A, B);
}

/* fast_corner_detect_10 + fast_corner_score_10 + fast_nonmax_3x3: row-major list of surviving corners */
int tbo_fast10(const u8 *img, int w, int h, int stride, int barrier, int32_t *xy, int32_t *scores, int cap)
{
    if (w < 7 || h < 7) return 0;
    int *sc = (int *)calloc((size_t)w * h, sizeof(int)); /* score+1 for corners, 0 otherwise */
    for (int y = 3; y < h - 3; y++)
        for (int x = 3; x < w - 3; x++) {
            int m = fast10_m(img + (size_t)y * stride + x, stride);
            if (m > barrier) sc[(size_t)y * w + x] = m; /* score = m - 1 >= barrier; keep m so that 0 = no corner */
        }
    int n = 0;
    for (int y = 3; y < h - 3; y++)
        for (int x = 3; x < w - 3; x++) {
            int m = sc[(size_t)y * w + x];
            if (!m) continue;
            int keep = 1;
            for (int dy = -1; dy <= 1 && keep; dy++)
                for (int dx = -1; dx <= 1; dx++) {
                    if (!dx && !dy) continue;
                    int m2 = sc[(size_t)(y + dy) * w + x + dx];
                    if (m2 && m2 >= m) { keep = 0; break; } /* Compare(X,Y) ((X)>=(Y)) */
                }
            if (!keep) continue;
            if (n < cap) { xy[2 * n] = x; xy[2 * n + 1] = y; scores[n] = m - 1; }
            n++;
        }
    free(sc);
    return n;
}

int tbo_fast_extract(const u8 *const *levels, const int *w, const int *h, const int *stride, int n_levels,
                     const float *inv_sf, int nFeatures, float threshold, const u8 *occupied, tbo_kp *out, int cap)
{
    if (n_levels <= 0 || !levels[0] || w[0] <= 0 || h[0] <= 0 || nFeatures <= 0) return 0;
    const int cell_size = (int)sqrtf((float)w[0] * (float)h[0] / (float)nFeatures);
    if (cell_size <= 0) return -1;
    const int grid_n_cols = (int)((float)w[0] / (float)cell_size);
    const int grid_n_rows = (int)((float)h[0] / (float)cell_size);
    const int ncell = grid_n_cols * grid_n_rows;
    tbo_kp *grid = (tbo_kp *)calloc((size_t)nFeatures, sizeof(tbo_kp)); /* gridPoints.resize(nFeatures): response 0 */
    int err = 0;
    for (int L = 0; L < n_levels && !err; L++) {
        int ccap = w[L] * h[L] / 2 + 16;
        int32_t *xy = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)ccap), *sc = (int32_t *)malloc(sizeof(int32_t) * (size_t)ccap);
        int n = tbo_fast10(levels[L], w[L], h[L], stride[L], 20, xy, sc, ccap);
        for (int i = 0; i < n; i++) {
            const int x = xy[2 * i], y = xy[2 * i + 1];
            const int k = (int)(((float)y * inv_sf[L]) / (float)cell_size) * grid_n_cols + (int)(((float)x * inv_sf[L]) / (float)cell_size);
            if (occupied && k >= 0 && k < ncell && occupied[k]) continue;
            const float score = tbo_shi_tomasi(levels[L], w[L], h[L], stride[L], x, y);
            if (k < 0 || k >= nFeatures) { err = 1; break; } /* gridPoints.at(k) throws */
            if (score > grid[k].response) {
                tbo_kp kp = {(float)x * inv_sf[L], (float)y * inv_sf[L], 1.f, 0.f, score, L, -1};
                grid[k] = kp;
            }
        }
        free(xy); free(sc);
    }
    int n = 0;
    if (!err)
        for (int k = 0; k < nFeatures; k++)
            if (grid[k].response > (float)threshold) { if (n < cap) out[n] = grid[k]; n++; }
    free(grid);
    return err ? -1 : n;
}

/* ORB stereo row-band matcher — block-commented body of LocalBA::AddMapPointsByStereo, LocalBA.cpp:69-288 (see header). */
int tbo_stereo_rowband(const tbo_stereo_params *p, const tbo_kp *kL, const u8 *dL, int nL,
                       const tbo_kp *kR, const u8 *dR, int nR, const float *inv_sf, int n_levels, int n_rows,
                       tbo_match *out, float *depth)
{
    /* row table (:78-93): vRowIndices[yi] lists right keys in push (= index) order */
    int *cnt = (int *)calloc((size_t)n_rows + 1, sizeof(int));
    int *lo = (int *)malloc(sizeof(int) * (size_t)(nR + 1)), *hi = (int *)malloc(sizeof(int) * (size_t)(nR + 1));
    for (int iR = 0; iR < nR; iR++) {
        const float kpY = kR[iR].y;
        int oc = kR[iR].octave;
        if (oc < 0) oc = 0;
        if (oc >= n_levels) oc = n_levels - 1;
        const float r = p->band * inv_sf[oc];
        const float a = kpY + r, b = kpY - r;
        int max_r = (int)ceilf(a), min_r = (int)floorf(b);
        if (min_r < 0) min_r = 0;
        if (max_r > n_rows - 1) max_r = n_rows - 1;
        lo[iR] = min_r; hi[iR] = max_r;
        for (int yi = min_r; yi <= max_r; yi++) cnt[yi + 1]++;
    }
    for (int y = 0; y < n_rows; y++) cnt[y + 1] += cnt[y];
    int *items = (int *)malloc(sizeof(int) * (size_t)(cnt[n_rows] + 1));
    int *cur = (int *)malloc(sizeof(int) * (size_t)(n_rows + 1));
    memcpy(cur, cnt, sizeof(int) * (size_t)(n_rows + 1));
    for (int iR = 0; iR < nR; iR++)
        for (int yi = lo[iR]; yi <= hi[iR]; yi++) items[cur[yi]++] = iR;
    int n = 0;
    for (int iL = 0; iL < nL; iL++) {
        depth[iL] = -1.0f;
        const int levelL = kL[iL].octave;
        const float vL = kL[iL].y, uL = kL[iL].x;
        if (!(vL > -1.f) || vL >= (float)n_rows) continue;
        const int row = (int)vL; /* (size_t)vL, :110 */
        int bestDist0 = 0x7fffffff, bestIdxR0 = 0;
        for (int k = cnt[row]; k < cnt[row + 1]; k++) {
            const int iR = items[k];
            if (kR[iR].octave < levelL - p->octave_tol || kR[iR].octave > levelL + p->octave_tol) continue; /* :131 */
            const float uR = kR[iR].x;
            if (uR >= p->min_u && uR <= p->max_u) { /* :136 */
                const int dist = tbo_hamming256(dL + 32 * (size_t)iL, dR + 32 * (size_t)iR);
                if (dist < bestDist0 && dist > 0) { bestDist0 = dist; bestIdxR0 = iR; } /* :140 */
            }
        }
        if (bestDist0 < p->th_dist) { /* :148 */
            out[n].queryIdx = iL; out[n].trainIdx = bestIdxR0; out[n].imgIdx = -1; out[n].distance = (float)bestDist0;
            n++;
            const float dx = fabsf(uL - kR[bestIdxR0].x); /* :168-171 */
            depth[iL] = dx > 0.f ? p->bf / dx : -1.0f;
        }
    }
    free(cnt); free(lo); free(hi); free(items); free(cur);
    return n;
}

/* ---- cv::pyrDown 8UC1 (imgproc/pyramids.cpp): 1-4-6-4-1 both ways in int, (v + 128) >> 8, BORDER_REFLECT_101 ---- */
void tbo_pyrdown_u8(const u8 *src, int w, int h, int sstride, u8 *dst, int dstride)
{
    const int dw = (w + 1) / 2, dh = (h + 1) / 2;
    int *rows = (int *)malloc(sizeof(int) * (size_t)dw * 5);
    for (int y = 0; y < dh; y++) {
        for (int k = 0; k < 5; k++) {
            const u8 *r = src + (size_t)reflect101(2 * y - 2 + k, h) * sstride;
            int *o = rows + (size_t)k * dw;
            for (int x = 0; x < dw; x++)
                o[x] = r[reflect101(2 * x, w)] * 6 + (r[reflect101(2 * x - 1, w)] + r[reflect101(2 * x + 1, w)]) * 4 +
                       r[reflect101(2 * x - 2, w)] + r[reflect101(2 * x + 2, w)];
        }
        for (int x = 0; x < dw; x++) {
            const int v = rows[2 * dw + x] * 6 + (rows[dw + x] + rows[3 * dw + x]) * 4 + rows[x] + rows[4 * dw + x];
            dst[(size_t)y * dstride + x] = (u8)((v + 128) >> 8);
        }
    }
    free(rows);
}

/* ---- cv::calcOpticalFlowPyrLK (video/lkpyramid.cpp), 8UC1, flags 0 ---- */
typedef struct { const u8 *p; int w, h, stride; } lk_img;
static inline int lk_px(const lk_img *im, int x, int y) /* pyramid level with its REFLECT_101 border */
{
    return im->p[(size_t)reflect101(y, im->h) * im->stride + reflect101(x, im->w)];
}
/* calcSharrDeriv at (x,y) inside the level; the derivative image has a zero (BORDER_CONSTANT) border */
static inline void lk_scharr(const lk_img *im, int x, int y, int *dx, int *dy)
{
    if (x < 0 || y < 0 || x >= im->w || y >= im->h) { *dx = *dy = 0; return; }
    int t0[3], t1[3];
    for (int k = 0; k < 3; k++) {
        const int xx = x + k - 1;
        const int a = lk_px(im, xx, y - 1), b = lk_px(im, xx, y), c = lk_px(im, xx, y + 1);
        t0[k] = (a + c) * 3 + b * 10;
        t1[k] = c - a;
    }
    *dx = t0[2] - t0[0];
    *dy = (t1[2] + t1[0]) * 3 + t1[1] * 10;
}
#define LK_DESCALE(x, n) (((x) + (1 << ((n) - 1))) >> (n))

int tbo_lk_flow(const tbo_lk_params *p, const u8 *prev, const u8 *next, int w, int h, int stride,
                const float *prev_xy, int n, float *next_xy, u8 *status, float *err)
{
    const int win = p->win;
    lk_img I[16], J[16];
    u8 *own[32];
    int n_own = 0, maxLevel = 0;
    I[0].p = prev; I[0].w = w; I[0].h = h; I[0].stride = stride;
    J[0].p = next; J[0].w = w; J[0].h = h; J[0].stride = stride;
    for (int l = 1; l <= p->max_level && l < 16; l++) { /* buildOpticalFlowPyramid */
        const int lw = (I[l - 1].w + 1) / 2, lh = (I[l - 1].h + 1) / 2;
        if (lw <= win || lh <= win) break;
        for (int k = 0; k < 2; k++) {
            lk_img *P = k ? J : I;
            u8 *b = (u8 *)malloc((size_t)lw * lh);
            tbo_pyrdown_u8(P[l - 1].p, P[l - 1].w, P[l - 1].h, P[l - 1].stride, b, lw);
            P[l].p = b; P[l].w = lw; P[l].h = lh; P[l].stride = lw;
            own[n_own++] = b;
        }
        maxLevel = l;
    }
    int maxCount = p->max_iter;
    if (maxCount < 0) maxCount = 0;
    if (maxCount > 100) maxCount = 100;
    double epsilon = (double)p->eps;
    if (epsilon < 0.) epsilon = 0.;
    if (epsilon > 10.) epsilon = 10.;
    epsilon *= epsilon;
    const float halfWin = (float)(win - 1) * 0.5f;
    const float FLT_SCALE = 1.f / (1 << 20);
    short *Ibuf = (short *)malloc(sizeof(short) * (size_t)win * win * 3);
    short *dIx = Ibuf + win * win, *dIy = dIx + win * win;
    for (int i = 0; i < n; i++) { status[i] = 1; if (err) err[i] = 0.f; }
    for (int level = maxLevel; level >= 0; level--) {
        const lk_img *Il = &I[level], *Jl = &J[level];
        for (int i = 0; i < n; i++) {
            const float sc = (float)(1. / (1 << level));
            float px = prev_xy[2 * i] * sc, py = prev_xy[2 * i + 1] * sc, nx, ny;
            if (level == maxLevel) { nx = px; ny = py; }
            else { nx = next_xy[2 * i] * 2.f; ny = next_xy[2 * i + 1] * 2.f; }
            next_xy[2 * i] = nx; next_xy[2 * i + 1] = ny;
            px -= halfWin; py -= halfWin;
            const int ipx = cv_floor_f(px), ipy = cv_floor_f(py);
            if (ipx < -win || ipx >= Il->w || ipy < -win || ipy >= Il->h) {
                if (level == 0) { status[i] = 0; if (err) err[i] = 0.f; }
                continue;
            }
            float a = px - (float)ipx, b = py - (float)ipy;
            int iw00 = cv_round_f((1.f - a) * (1.f - b) * (float)(1 << 14)), iw01 = cv_round_f(a * (1.f - b) * (float)(1 << 14)),
                iw10 = cv_round_f((1.f - a) * b * (float)(1 << 14)), iw11 = (1 << 14) - iw00 - iw01 - iw10;
            int64_t iA11 = 0, iA12 = 0, iA22 = 0;
            for (int y = 0; y < win; y++)
                for (int x = 0; x < win; x++) {
                    const int X = ipx + x, Y = ipy + y;
                    int dx00, dy00, dx01, dy01, dx10, dy10, dx11, dy11;
                    lk_scharr(Il, X, Y, &dx00, &dy00); lk_scharr(Il, X + 1, Y, &dx01, &dy01);
                    lk_scharr(Il, X, Y + 1, &dx10, &dy10); lk_scharr(Il, X + 1, Y + 1, &dx11, &dy11);
                    const int ival = LK_DESCALE(lk_px(Il, X, Y) * iw00 + lk_px(Il, X + 1, Y) * iw01 + lk_px(Il, X, Y + 1) * iw10 + lk_px(Il, X + 1, Y + 1) * iw11, 9);
                    const int ixval = LK_DESCALE(dx00 * iw00 + dx01 * iw01 + dx10 * iw10 + dx11 * iw11, 14);
                    const int iyval = LK_DESCALE(dy00 * iw00 + dy01 * iw01 + dy10 * iw10 + dy11 * iw11, 14);
                    Ibuf[y * win + x] = (short)ival; dIx[y * win + x] = (short)ixval; dIy[y * win + x] = (short)iyval;
                    iA11 += (int64_t)ixval * ixval; iA12 += (int64_t)ixval * iyval; iA22 += (int64_t)iyval * iyval;
                }
            const float A11 = (float)iA11 * FLT_SCALE, A12 = (float)iA12 * FLT_SCALE, A22 = (float)iA22 * FLT_SCALE;
            float D = A11 * A22 - A12 * A12;
            const float minEig = (A22 + A11 - sqrtf((A11 - A22) * (A11 - A22) + 4.f * A12 * A12)) / (float)(2 * win * win);
            if ((double)minEig < (double)p->min_eig || D < 1.1920929e-07f) {
                if (level == 0) status[i] = 0;
                continue;
            }
            D = 1.f / D;
            nx -= halfWin; ny -= halfWin;
            float pdx = 0.f, pdy = 0.f;
            for (int j = 0; j < maxCount; j++) {
                const int inx = cv_floor_f(nx), iny = cv_floor_f(ny);
                if (inx < -win || inx >= Jl->w || iny < -win || iny >= Jl->h) {
                    if (level == 0) status[i] = 0;
                    break;
                }
                a = nx - (float)inx; b = ny - (float)iny;
                iw00 = cv_round_f((1.f - a) * (1.f - b) * (float)(1 << 14)); iw01 = cv_round_f(a * (1.f - b) * (float)(1 << 14));
                iw10 = cv_round_f((1.f - a) * b * (float)(1 << 14)); iw11 = (1 << 14) - iw00 - iw01 - iw10;
                int64_t ib1 = 0, ib2 = 0;
                for (int y = 0; y < win; y++)
                    for (int x = 0; x < win; x++) {
                        const int X = inx + x, Y = iny + y;
                        const int diff = LK_DESCALE(lk_px(Jl, X, Y) * iw00 + lk_px(Jl, X + 1, Y) * iw01 + lk_px(Jl, X, Y + 1) * iw10 + lk_px(Jl, X + 1, Y + 1) * iw11, 9) - Ibuf[y * win + x];
                        ib1 += (int64_t)diff * dIx[y * win + x]; ib2 += (int64_t)diff * dIy[y * win + x];
                    }
                const float b1 = (float)ib1 * FLT_SCALE, b2 = (float)ib2 * FLT_SCALE;
                const float ddx = (A12 * b2 - A22 * b1) * D, ddy = (A12 * b1 - A11 * b2) * D;
                nx += ddx; ny += ddy;
                next_xy[2 * i] = nx + halfWin; next_xy[2 * i + 1] = ny + halfWin;
                if ((double)ddx * (double)ddx + (double)ddy * (double)ddy <= epsilon) break;
                if (j > 0 && fabs((double)(ddx + pdx)) < 0.01 && fabs((double)(ddy + pdy)) < 0.01) {
                    next_xy[2 * i] -= ddx * 0.5f; next_xy[2 * i + 1] -= ddy * 0.5f;
                    break;
                }
                pdx = ddx; pdy = ddy;
            }
            if (status[i] && level == 0) { /* the reference passes `err`, so the final window check runs */
                const float fx = next_xy[2 * i] - halfWin, fy = next_xy[2 * i + 1] - halfWin;
                const int inx = cv_floor_f(fx), iny = cv_floor_f(fy);
                if (inx < -win || inx >= Jl->w || iny < -win || iny >= Jl->h) { status[i] = 0; continue; }
                a = fx - (float)inx; b = fy - (float)iny;
                iw00 = cv_round_f((1.f - a) * (1.f - b) * (float)(1 << 14)); iw01 = cv_round_f(a * (1.f - b) * (float)(1 << 14));
                iw10 = cv_round_f((1.f - a) * b * (float)(1 << 14)); iw11 = (1 << 14) - iw00 - iw01 - iw10;
                float errval = 0.f;
                for (int y = 0; y < win; y++)
                    for (int x = 0; x < win; x++) {
                        const int X = inx + x, Y = iny + y;
                        const int diff = LK_DESCALE(lk_px(Jl, X, Y) * iw00 + lk_px(Jl, X + 1, Y) * iw01 + lk_px(Jl, X, Y + 1) * iw10 + lk_px(Jl, X + 1, Y + 1) * iw11, 9) - Ibuf[y * win + x];
                        errval += fabsf((float)diff);
                    }
                if (err) err[i] = errval * 1.f / (float)(32 * win * win);
            }
        }
    }
    free(Ibuf);
    for (int k = 0; k < n_own; k++) free(own[k]);
    return maxLevel;
}

/* PinholeCamera::UndistortPoint — CameraModel.cpp:104-121; cv::undistortPoints (calib3d/undistort) with the default
 * TermCriteria(MAX_ITER, 5, 0.01), R = I, P = K; K and D are float matrices promoted to double. */
void tbo_undistort_points(float ffx, float ffy, float fcx, float fcy, const float d[5], const float *xy_in, int n, float *xy_out)
{
    if (!(fabsf(d[0]) > 1e-7f)) { /* mbDistortion false: returned as is */
        memcpy(xy_out, xy_in, sizeof(float) * 2 * (size_t)n);
        return;
    }
    const double fx = ffx, fy = ffy, cx = fcx, cy = fcy;
    const double k0 = d[0], k1 = d[1], k2 = d[2], k3 = d[3], k4 = d[4];
    const double ifx = 1. / fx, ify = 1. / fy;
    for (int i = 0; i < n; i++) {
        const double u = xy_in[2 * i], v = xy_in[2 * i + 1];
        double x = (u - cx) * ifx, y = (v - cy) * ify;
        const double x0 = x, y0 = y;
        for (int j = 0; j < 5; j++) {
            const double r2 = x * x + y * y;
            const double icdist = (1 + ((0. * r2 + 0.) * r2 + 0.) * r2) / (1 + ((k4 * r2 + k1) * r2 + k0) * r2);
            if (icdist < 0) { x = (u - cx) * ifx; y = (v - cy) * ify; break; }
            const double deltaX = 2 * k2 * x * y + k3 * (r2 + 2 * x * x) + 0. * r2 + 0. * r2 * r2;
            const double deltaY = k2 * (r2 + 2 * y * y) + 2 * k3 * x * y + 0. * r2 + 0. * r2 * r2;
            x = (x0 - deltaX) * icdist;
            y = (y0 - deltaY) * icdist;
        }
        const double xx = fx * x + 0. * y + cx, yy = 0. * x + fy * y + cy, ww = 1. / (0. * x + 0. * y + 1.);
        xy_out[2 * i] = (float)(xx * ww);
        xy_out[2 * i + 1] = (float)(yy * ww);
    }
}


/* ------------------------------------------------------------------------------------------------------------------
 * Matcher::rejectWithF (matcher.cpp:853-881) -> cv::findFundamentalMat(FM_RANSAC, 1.0, 0.99): OpenCV 4.x restated
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct { uint64_t state; } tbo_rng;                       /* cv::RNG: multiply-with-carry */
static unsigned rng_next(tbo_rng *r) { r->state = (uint64_t)(unsigned)r->state * 4164903690u + (unsigned)(r->state >> 32); return (unsigned)r->state; }
static int rng_uniform(tbo_rng *r, int a, int b) { return a == b ? a : (int)(rng_next(r) % (unsigned)(b - a) + a); }

/* haveCollinearPoints(m, 7): the last point against every pair of earlier ones */
static int fm_collinear7(const float *p)
{
    const int i = 6;
    for (int j = 0; j < i; j++) {
        double dx1 = (double)p[2 * j] - p[2 * i], dy1 = (double)p[2 * j + 1] - p[2 * i + 1];
        for (int k = 0; k < j; k++) {
            double dx2 = (double)p[2 * k] - p[2 * i], dy2 = (double)p[2 * k + 1] - p[2 * i + 1];
            if (fabs(dx2 * dy1 - dy2 * dx1) <= (double)FLT_EPSILON * (fabs(dx1) + fabs(dy1) + fabs(dx2) + fabs(dy2))) return 1;
        }
    }
    return 0;
}
/* RANSACPointSetRegistrator::getSubset (modelPoints = 7, maxAttempts = 10000) */
static int fm_get_subset(const float *m1, const float *m2, int count, tbo_rng *rng, float *s1, float *s2)
{
    for (int it = 0; it < 10000; it++) {
        int idx[7];
        for (int i = 0; i < 7; i++) {
            int v, dup;
            do {
                v = rng_uniform(rng, 0, count); dup = 0;
                for (int k = 0; k < i; k++) dup |= idx[k] == v;
            } while (dup);
            idx[i] = v;
            s1[2 * i] = m1[2 * v]; s1[2 * i + 1] = m1[2 * v + 1]; s2[2 * i] = m2[2 * v]; s2[2 * i + 1] = m2[2 * v + 1];
        }
        if (!fm_collinear7(s1) && !fm_collinear7(s2)) return 1;
    }
    return 0;
}
/* cv::solveCubic for c[0] x^3 + c[1] x^2 + c[2] x + c[3] */
static int fm_solve_cubic(const double *c, double *x)
{
    double a0 = c[0], a1 = c[1], a2 = c[2], a3 = c[3];
    if (a0 == 0) {
        if (a1 == 0) { if (a2 == 0) return 0; x[0] = -a3 / a2; return 1; }
        double d = a2 * a2 - 4 * a1 * a3;
        if (d < 0) return 0;
        d = sqrt(d);
        double q1 = (-a2 + d) * 0.5, q2 = (a2 + d) * -0.5;
        if (fabs(q1) > fabs(q2)) { x[0] = q1 / a1; x[1] = a3 / q1; } else { x[0] = q2 / a1; x[1] = a3 / q2; }
        return d > 0 ? 2 : 1;
    }
    a0 = 1. / a0; a1 *= a0; a2 *= a0; a3 *= a0;
    double Q = (a1 * a1 - 3 * a2) * (1. / 9), R = (2 * a1 * a1 * a1 - 9 * a1 * a2 + 27 * a3) * (1. / 54);
    double Qcubed = Q * Q * Q, d = Qcubed - R * R;
    if (d > 0) {
        double theta = acos(R / sqrt(Qcubed)), sqrtQ = sqrt(Q), t0 = -2 * sqrtQ, t1 = theta * (1. / 3), t2 = a1 * (1. / 3);
        x[0] = t0 * cos(t1) - t2; x[1] = t0 * cos(t1 + (2. * 3.1415926535897932384626433832795 / 3)) - t2;
        x[2] = t0 * cos(t1 + (4. * 3.1415926535897932384626433832795 / 3)) - t2;
        return 3;
    }
    if (d == 0) {
        if (R >= 0) { x[0] = -2 * pow(R, 1. / 3) - a1 / 3; x[1] = pow(R, 1. / 3) - a1 / 3; }
        else { x[0] = 2 * pow(-R, 1. / 3) - a1 / 3; x[1] = -pow(-R, 1. / 3) - a1 / 3; }
        return x[0] == x[1] ? 1 : 2;
    }
    d = sqrt(-d);
    double e = pow(d + fabs(R), 1. / 3);
    if (R > 0) e = -e;
    x[0] = (e + Q / e) - a1 * (1. / 3);
    return 1;
}
/* the two eigenvectors of the smallest eigenvalues of the symmetric 9x9 M (cyclic Jacobi) */
static void fm_null2(double M[9][9], double *f1, double *f2)
{
    double V[9][9];
    for (int i = 0; i < 9; i++) for (int j = 0; j < 9; j++) V[i][j] = i == j;
    for (int sweep = 0; sweep < 60; sweep++) {
        double off = 0;
        for (int p = 0; p < 9; p++) for (int q = p + 1; q < 9; q++) off += M[p][q] * M[p][q];
        if (off < 1e-300) break;
        for (int p = 0; p < 9; p++) for (int q = p + 1; q < 9; q++) {
            if (fabs(M[p][q]) < 1e-300) continue;
            double theta = (M[q][q] - M[p][p]) / (2 * M[p][q]);
            double t = (theta >= 0 ? 1. : -1.) / (fabs(theta) + sqrt(theta * theta + 1));
            double c = 1 / sqrt(t * t + 1), s = t * c;
            for (int k = 0; k < 9; k++) { double a = M[k][p], b = M[k][q]; M[k][p] = c * a - s * b; M[k][q] = s * a + c * b; }
            for (int k = 0; k < 9; k++) { double a = M[p][k], b = M[q][k]; M[p][k] = c * a - s * b; M[q][k] = s * a + c * b; }
            for (int k = 0; k < 9; k++) { double a = V[k][p], b = V[k][q]; V[k][p] = c * a - s * b; V[k][q] = s * a + c * b; }
        }
    }
    int i1 = 0, i2 = -1;
    for (int i = 1; i < 9; i++) if (M[i][i] < M[i1][i1]) i1 = i;
    for (int i = 0; i < 9; i++) if (i != i1 && (i2 < 0 || M[i][i] < M[i2][i2])) i2 = i;
    for (int k = 0; k < 9; k++) { f1[k] = V[k][i2]; f2[k] = V[k][i1]; }   /* OpenCV: f1 = 8th, f2 = 9th (smallest) right singular vector */
}
/* run7Point (fundam.cpp): up to 3 models, each 9 doubles */
static int fm_run7(const float *m1, const float *m2, double *F)
{
    double c1x = 0, c1y = 0, c2x = 0, c2y = 0, scale1 = 0, scale2 = 0;
    for (int i = 0; i < 7; i++) { c1x += m1[2 * i]; c1y += m1[2 * i + 1]; c2x += m2[2 * i]; c2y += m2[2 * i + 1]; }
    const double t = 1. / 7;
    c1x *= t; c1y *= t; c2x *= t; c2y *= t;
    for (int i = 0; i < 7; i++) {
        double dx = m1[2 * i] - c1x, dy = m1[2 * i + 1] - c1y; scale1 += sqrt(dx * dx + dy * dy);
        dx = m2[2 * i] - c2x; dy = m2[2 * i + 1] - c2y; scale2 += sqrt(dx * dx + dy * dy);
    }
    scale1 *= t; scale2 *= t;
    if (scale1 < FLT_EPSILON || scale2 < FLT_EPSILON) return 0;
    scale1 = sqrt(2.) / scale1; scale2 = sqrt(2.) / scale2;
    double A[7][9], M[9][9];
    for (int i = 0; i < 7; i++) {
        double x0 = (m1[2 * i] - c1x) * scale1, y0 = (m1[2 * i + 1] - c1y) * scale1;
        double x1 = (m2[2 * i] - c2x) * scale2, y1 = (m2[2 * i + 1] - c2y) * scale2;
        double r[9] = {x1 * x0, x1 * y0, x1, y1 * x0, y1 * y0, y1, x0, y0, 1};
        for (int k = 0; k < 9; k++) A[i][k] = r[k];
    }
    for (int a = 0; a < 9; a++) for (int b = 0; b < 9; b++) { double s = 0; for (int i = 0; i < 7; i++) s += A[i][a] * A[i][b]; M[a][b] = s; }
    double f1[9], f2[9], c[4], r[3] = {0, 0, 0};
    fm_null2(M, f1, f2);
    for (int i = 0; i < 9; i++) f1[i] -= f2[i];
    double t0 = f2[4] * f2[8] - f2[5] * f2[7], t1 = f2[3] * f2[8] - f2[5] * f2[6], t2 = f2[3] * f2[7] - f2[4] * f2[6];
    c[3] = f2[0] * t0 - f2[1] * t1 + f2[2] * t2;
    c[2] = f1[0] * t0 - f1[1] * t1 + f1[2] * t2 - f1[3] * (f2[1] * f2[8] - f2[2] * f2[7]) + f1[4] * (f2[0] * f2[8] - f2[2] * f2[6]) -
           f1[5] * (f2[0] * f2[7] - f2[1] * f2[6]) + f1[6] * (f2[1] * f2[5] - f2[2] * f2[4]) - f1[7] * (f2[0] * f2[5] - f2[2] * f2[3]) +
           f1[8] * (f2[0] * f2[4] - f2[1] * f2[3]);
    t0 = f1[4] * f1[8] - f1[5] * f1[7]; t1 = f1[3] * f1[8] - f1[5] * f1[6]; t2 = f1[3] * f1[7] - f1[4] * f1[6];
    c[1] = f2[0] * t0 - f2[1] * t1 + f2[2] * t2 - f2[3] * (f1[1] * f1[8] - f1[2] * f1[7]) + f2[4] * (f1[0] * f1[8] - f1[2] * f1[6]) -
           f2[5] * (f1[0] * f1[7] - f1[1] * f1[6]) + f2[6] * (f1[1] * f1[5] - f1[2] * f1[4]) - f2[7] * (f1[0] * f1[5] - f1[2] * f1[3]) +
           f2[8] * (f1[0] * f1[4] - f1[1] * f1[3]);
    c[0] = f1[0] * t0 - f1[1] * t1 + f1[2] * t2;
    const int n = fm_solve_cubic(c, r);
    if (n < 1 || n > 3) return 0;
    const double T1[9] = {scale1, 0, -scale1 * c1x, 0, scale1, -scale1 * c1y, 0, 0, 1};
    const double T2[9] = {scale2, 0, -scale2 * c2x, 0, scale2, -scale2 * c2y, 0, 0, 1};
    for (int k = 0; k < n; k++, F += 9) {
        double lambda = r[k], mu = 1., s = f1[8] * r[k] + f2[8], G[9], H[9];
        if (fabs(s) > DBL_EPSILON) { mu = 1. / s; lambda *= mu; G[8] = 1.; } else G[8] = 0.;
        for (int i = 0; i < 8; i++) G[i] = f1[i] * lambda + f2[i] * mu;
        /* F = T2^T * G * T1 */
        for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) { double v = 0; for (int m = 0; m < 3; m++) v += T2[m * 3 + a] * G[m * 3 + b]; H[a * 3 + b] = v; }
        for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) { double v = 0; for (int m = 0; m < 3; m++) v += H[a * 3 + m] * T1[m * 3 + b]; F[a * 3 + b] = v; }
        if (fabs(F[8]) > FLT_EPSILON) { const double inv = 1. / F[8]; for (int i = 0; i < 9; i++) F[i] *= inv; }
    }
    return n;
}
/* FMEstimatorCallback::computeError + findInliers: err <= threshold^2 (float compare) */
static int fm_inliers(const float *m1, const float *m2, int count, const double *F, float thr2, uint8_t *mask)
{
    int good = 0;
    for (int i = 0; i < count; i++) {
        const double x1 = m1[2 * i], y1 = m1[2 * i + 1], x2 = m2[2 * i], y2 = m2[2 * i + 1];
        double a = F[0] * x1 + F[1] * y1 + F[2], b = F[3] * x1 + F[4] * y1 + F[5], c = F[6] * x1 + F[7] * y1 + F[8];
        const double s2 = 1. / (a * a + b * b), d2 = x2 * a + y2 * b + c;
        a = F[0] * x2 + F[3] * y2 + F[6]; b = F[1] * x2 + F[4] * y2 + F[7]; c = F[2] * x2 + F[5] * y2 + F[8];
        const double s1 = 1. / (a * a + b * b), d1 = x1 * a + y1 * b + c;
        const double e1 = d1 * d1 * s1, e2 = d2 * d2 * s2;
        const float err = (float)(e1 > e2 ? e1 : e2);
        const int in = err <= thr2;
        if (mask) mask[i] = (uint8_t)in;
        good += in;
    }
    return good;
}
static int fm_update_iters(double p, double ep, int model_points, int max_iters)
{
    p = p < 0 ? 0 : p > 1 ? 1 : p; ep = ep < 0 ? 0 : ep > 1 ? 1 : ep;
    double num = 1. - p > DBL_MIN ? 1. - p : DBL_MIN, denom = 1. - pow(1. - ep, model_points);
    if (denom < DBL_MIN) return 0;
    num = log(num); denom = log(denom);
    return denom >= 0 || -num >= max_iters * (-denom) ? max_iters : (int)lrint(num / denom);
}
int tbo_reject_with_f(const float *cur_xy, const float *last_xy, int n, u8 *status, double threshold, double confidence,
                      int max_iters, double *F9)
{
    if (n <= 8) return 0;                                         /* matcher.cpp:870: findFundamentalMat is not called */
    int *id = (int *)malloc(sizeof(int) * (size_t)n), count = 0;
    float *m1 = (float *)malloc(sizeof(float) * 2 * (size_t)n), *m2 = (float *)malloc(sizeof(float) * 2 * (size_t)n);
    for (int i = 0; i < n; i++) if (status[i]) {
        id[count] = i; m1[2 * count] = cur_xy[2 * i]; m1[2 * count + 1] = cur_xy[2 * i + 1];
        m2[2 * count] = last_xy[2 * i]; m2[2 * count + 1] = last_xy[2 * i + 1]; count++;
    }
    int iters = 0;
    if (count >= 15) {                                            /* below that OpenCV runs LMedS (or nothing at all) */
        if (threshold <= 0) threshold = 3;
        if (confidence < DBL_EPSILON || confidence > 1 - DBL_EPSILON) confidence = 0.99;
        u8 *mask = (u8 *)malloc((size_t)count);
        tbo_rng rng = {0xffffffffffffffffull};
        int niters = max_iters > 1 ? max_iters : 1, max_good = 0;
        double best[9] = {0}, models[27];
        const float thr2 = (float)(threshold * threshold);
        for (iters = 0; iters < niters; iters++) {
            float s1[14], s2[14];
            if (!fm_get_subset(m1, m2, count, &rng, s1, s2)) break;
            const int nm = fm_run7(s1, s2, models);
            for (int k = 0; k < nm; k++) {
                const int good = fm_inliers(m1, m2, count, models + 9 * k, thr2, NULL);
                if (good > (max_good > 6 ? max_good : 6)) {
                    memcpy(best, models + 9 * k, sizeof best);
                    max_good = good;
                    niters = fm_update_iters(confidence, (double)(count - good) / count, 7, niters);
                }
            }
        }
        if (max_good > 0) {
            fm_inliers(m1, m2, count, best, thr2, mask);
            for (int i = 0; i < count; i++) if (!mask[i]) status[id[i]] = 0;   /* matcher.cpp:874-880 */
            if (F9) memcpy(F9, best, sizeof best);
        }
        free(mask);
    }
    free(id); free(m1); free(m2);
    return iters;
}
