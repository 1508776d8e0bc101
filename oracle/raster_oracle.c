/* oracle/raster_oracle.c — CPU restatement of feh::Renderer's depth pass.  TEST INFRASTRUCTURE ONLY.
 *
 * PARITY UNPINNED BY THE REFERENCE: VISMA ships no golden depth map and its renderer needs an OpenGL
 * context + glm, neither of which exists in this container.  What is restated here, with sources:
 *   projection  render/renderer.cpp:238-268 (frustum extents from intrinsics, top/bottom flipped) and the
 *               published definition of glm::frustum (RH, z in [-1,1]); glm itself is not vendored.
 *   view        render/renderer.cpp:250-254, 284-300: diag(1,-1,-1,1) * pose
 *   vertex      render/shaders/basic_mvp.vert:10: gl_Position = projection * view * model * vec4(p, 1)
 *   state       render/renderer.cpp:321-351: depth test GL_LESS, clear depth 1, no culling, no fragment
 *               shader; 24-bit depth attachment (renderer.cpp:213-215); read-back as float (:343)
 *   GL 3.3 fixed function: clip to -w <= z <= w, perspective divide, viewport (0,0,W,H), depth range
 *               [0,1], pixel-centre sampling, window-space-linear z, unorm24 = round(z * (2^24-1)).
 * What GL leaves to the vendor and we pin canonically (the CUDA rasteriser implements the same rules and
 * must match bit for bit):
 *   - float32 vertex stage evaluated as MVP = (P*V)*M then MVP*[p,1], each dot product accumulated left
 *     to right with separately rounded multiplies and adds (no FMA);
 *   - everything after the vertex stage in IEEE double, no FMA: near/far polygon clipping, divide,
 *     viewport; x_w,y_w snapped to 1/256 pixel (round-half-even), clamped to +-2^29 sub-pixels;
 *   - int64 edge functions, top-left fill rule in image space (y down);
 *   - z(px) = ((E0*z0 + E1*z1) + E2*z2) / (E0+E1+E2) in double, q = rint(z*(2^24-1)) clamped to [0,2^24-1].
 * x/y frustum planes are not clipped geometrically (guard band): the pixel loop is bounded by the image.
 */
#include "oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define ZMAX24 16777215u
#define SUBPIX 256
#define COORD_CLAMP 536870912.0 /* 2^29 sub-pixels */

/* Renderer::SetCamera(zn, zf, intrinsics) (render/renderer.cpp:238-268).  Column-major like glm. */
void vo_projection(float zn, float zf, float fx, float fy, float cx, float cy, int H, int W, float P[16]) {
    float left = -cx / fx * zn;                                          /* :259 */
    float right = (float)(((double)(float)W - 1.0 - (double)cx) / (double)fx * (double)zn); /* :260 (the 1.0 literal promotes to double) */
    float bottom = cy / fy * zn;                                          /* :266 flipped */
    float top = (cy - (float)(H - 1)) / fy * zn;                          /* :267 */
    memset(P, 0, sizeof(float) * 16);
    /* glm::frustum(left, right, bottom, top, near, far), RH_NO */
    P[0] = (2.0f * zn) / (right - left);
    P[5] = (2.0f * zn) / (top - bottom);
    P[8] = (right + left) / (right - left);
    P[9] = (top + bottom) / (top - bottom);
    P[10] = -(zf + zn) / (zf - zn);
    P[11] = -1.0f;
    P[14] = -(2.0f * zf * zn) / (zf - zn);
}

/* c = a * b for column-major float 4x4, each element ((a0*b0 + a1*b1) + a2*b2) + a3*b3 */
static void mat4f_mul(const float a[16], const float b[16], float c[16]) {
    float r[16];
    for (int col = 0; col < 4; col++)
        for (int row = 0; row < 4; row++) {
            float s = a[0 * 4 + row] * b[col * 4 + 0];
            s = s + a[1 * 4 + row] * b[col * 4 + 1];
            s = s + a[2 * 4 + row] * b[col * 4 + 2];
            s = s + a[3 * 4 + row] * b[col * 4 + 3];
            r[col * 4 + row] = s;
        }
    memcpy(c, r, sizeof(r));
}

/* Renderer::SetCamera(pose) (render/renderer.cpp:284-300) */
void vo_view(const float pose[16], float V[16]) {
    const float v2g[16] = {1, 0, 0, 0, 0, -1, 0, 0, 0, 0, -1, 0, 0, 0, 0, 1};
    mat4f_mul(v2g, pose, V);
}

float vo_linearize_depth(float zb, float zn, float zf) { /* render/renderer.h:32-36 */
    return 2 * zn * zf / (zf + zn - (2 * zb - 1) * (zf - zn));
}

typedef struct { double x, y, z, w; } clipv;

static int clip_poly(const clipv *in, int n, clipv *out, int plane) {
    /* plane 0: near  d = w + z >= 0 ; plane 1: far  d = w - z >= 0 */
    int m = 0;
    for (int i = 0; i < n; i++) {
        const clipv *a = &in[i], *b = &in[(i + 1) % n];
        double da = plane == 0 ? a->w + a->z : a->w - a->z;
        double db = plane == 0 ? b->w + b->z : b->w - b->z;
        int ia = da >= 0.0, ib = db >= 0.0;
        if (ia) out[m++] = *a;
        if (ia != ib) {
            double t = da / (da - db);
            clipv c;
            c.x = a->x + t * (b->x - a->x);
            c.y = a->y + t * (b->y - a->y);
            c.z = a->z + t * (b->z - a->z);
            c.w = a->w + t * (b->w - a->w);
            out[m++] = c;
        }
    }
    return m;
}

static int64_t snap(double v) {
    double s = v * (double)SUBPIX;
    if (!(s > -COORD_CLAMP)) s = -COORD_CLAMP; /* also catches NaN */
    if (s > COORD_CLAMP) s = COORD_CLAMP;
    return (int64_t)llrint(s);
}

static int64_t floor_div(int64_t a, int64_t b) { /* b > 0 */
    int64_t q = a / b, r = a % b;
    return (r != 0 && r < 0) ? q - 1 : q;
}

static int top_left(int64_t dx, int64_t dy) { return dy < 0 || (dy == 0 && dx > 0); }

static void raster_tri(int64_t X0, int64_t Y0, double z0, int64_t X1, int64_t Y1, double z1, int64_t X2,
                       int64_t Y2, double z2, int H, int W, uint32_t *zbuf) {
    int64_t area2 = (X1 - X0) * (Y2 - Y0) - (X2 - X0) * (Y1 - Y0);
    if (area2 == 0) return;
    if (area2 < 0) { /* make the orientation positive; culling is off so both faces draw */
        int64_t t;
        double tz;
        t = X1; X1 = X2; X2 = t;
        t = Y1; Y1 = Y2; Y2 = t;
        tz = z1; z1 = z2; z2 = tz;
        area2 = -area2;
    }
    int64_t minX = X0 < X1 ? (X0 < X2 ? X0 : X2) : (X1 < X2 ? X1 : X2);
    int64_t maxX = X0 > X1 ? (X0 > X2 ? X0 : X2) : (X1 > X2 ? X1 : X2);
    int64_t minY = Y0 < Y1 ? (Y0 < Y2 ? Y0 : Y2) : (Y1 < Y2 ? Y1 : Y2);
    int64_t maxY = Y0 > Y1 ? (Y0 > Y2 ? Y0 : Y2) : (Y1 > Y2 ? Y1 : Y2);
    /* pixel i has its centre at 256*i + 128 */
    int64_t i0 = floor_div(minX - 128 + (SUBPIX - 1), SUBPIX), i1 = floor_div(maxX - 128, SUBPIX);
    int64_t j0 = floor_div(minY - 128 + (SUBPIX - 1), SUBPIX), j1 = floor_div(maxY - 128, SUBPIX);
    if (i0 < 0) i0 = 0;
    if (j0 < 0) j0 = 0;
    if (i1 > W - 1) i1 = W - 1;
    if (j1 > H - 1) j1 = H - 1;
    /* edge a->b: E(p) = (Xb-Xa)*(py-Ya) - (Yb-Ya)*(px-Xa); E0 is the edge opposite v0 (v1->v2) etc. */
    int64_t dx0 = X2 - X1, dy0 = Y2 - Y1, dx1 = X0 - X2, dy1 = Y0 - Y2, dx2 = X1 - X0, dy2 = Y1 - Y0;
    int tl0 = top_left(dx0, dy0), tl1 = top_left(dx1, dy1), tl2 = top_left(dx2, dy2);
    double a2 = (double)area2;
    for (int64_t j = j0; j <= j1; j++) {
        int64_t py = j * SUBPIX + 128;
        for (int64_t i = i0; i <= i1; i++) {
            int64_t px = i * SUBPIX + 128;
            int64_t E0 = dx0 * (py - Y1) - dy0 * (px - X1);
            int64_t E1 = dx1 * (py - Y2) - dy1 * (px - X2);
            int64_t E2 = dx2 * (py - Y0) - dy2 * (px - X0);
            if (!(E0 > 0 || (E0 == 0 && tl0))) continue;
            if (!(E1 > 0 || (E1 == 0 && tl1))) continue;
            if (!(E2 > 0 || (E2 == 0 && tl2))) continue;
            double z = (((double)E0 * z0 + (double)E1 * z1) + (double)E2 * z2) / a2;
            double qd = z * (double)ZMAX24;
            int64_t q = (int64_t)llrint(qd);
            if (!(qd > 0.0)) q = 0;
            if (q > (int64_t)ZMAX24) q = ZMAX24;
            uint32_t *dst = zbuf + j * W + i;
            if ((uint32_t)q < *dst) *dst = (uint32_t)q; /* GL_LESS */
        }
    }
}

int vo_render_depth(const float *V, int64_t nV, const int32_t *F, int64_t nF, const float model[16],
                    const float view[16], const float proj[16], int H, int W, uint32_t *out_z24,
                    float *out_depth) {
    if (H <= 0 || W <= 0) return -1;
    uint32_t *zbuf = out_z24 ? out_z24 : (uint32_t *)malloc(sizeof(uint32_t) * (size_t)H * W);
    for (int64_t p = 0; p < (int64_t)H * W; p++) zbuf[p] = ZMAX24; /* glClear depth = 1 */
    float PV[16], MVP[16];
    mat4f_mul(proj, view, PV);
    mat4f_mul(PV, model, MVP);
    for (int64_t f = 0; f < nF; f++) {
        clipv poly[8], tmp[8];
        int ok = 1;
        for (int k = 0; k < 3; k++) {
            int32_t vi = F[3 * f + k];
            if (vi < 0 || vi >= nV) { ok = 0; break; }
            float x = V[3 * vi], y = V[3 * vi + 1], z = V[3 * vi + 2];
            float c[4];
            for (int r = 0; r < 4; r++) {
                float s = MVP[0 * 4 + r] * x;
                s = s + MVP[1 * 4 + r] * y;
                s = s + MVP[2 * 4 + r] * z;
                s = s + MVP[3 * 4 + r];
                c[r] = s;
            }
            poly[k].x = c[0]; poly[k].y = c[1]; poly[k].z = c[2]; poly[k].w = c[3];
        }
        if (!ok) continue;
        int n = clip_poly(poly, 3, tmp, 0);
        if (n < 3) continue;
        n = clip_poly(tmp, n, poly, 1);
        if (n < 3) continue;
        int64_t X[8], Y[8];
        double Z[8];
        for (int k = 0; k < n; k++) {
            double iw = 1.0 / poly[k].w;
            double xn = poly[k].x * iw, yn = poly[k].y * iw, zn = poly[k].z * iw;
            X[k] = snap((xn + 1.0) * ((double)W * 0.5));
            Y[k] = snap((yn + 1.0) * ((double)H * 0.5));
            Z[k] = (zn + 1.0) * 0.5;
        }
        for (int k = 1; k + 1 < n; k++)
            raster_tri(X[0], Y[0], Z[0], X[k], Y[k], Z[k], X[k + 1], Y[k + 1], Z[k + 1], H, W, zbuf);
    }
    if (out_depth)
        for (int64_t p = 0; p < (int64_t)H * W; p++)
            out_depth[p] = (float)((double)zbuf[p] / (double)ZMAX24);
    if (!out_z24) free(zbuf);
    return 0;
}

/* ---- RenderEdge / RenderMask (render/renderer.cpp:353-433, render/shaders/edge_detection.frag:14-76) ----
 * RenderEdge = RenderDepth followed by a full-screen pass that linearises the depth texture with the shader's
 * OWN z_near / z_far uniforms (fixed to 0.05 / 2.0 at construction, renderer.cpp:95-96 — not the camera's),
 * takes the mean absolute difference of the four opposite neighbour pairs and soft-thresholds it to [0,1];
 * a 5-texel border and background pixels are 0.  The quad maps uv (0,0) to the lower-left corner
 * (renderer.cpp:147-153) and the colour attachment is read back as GL_UNSIGNED_BYTE (:392): unorm8 = rint(c*255).
 * PARITY UNPINNED (vendor texture filtering / float evaluation); float ops below are evaluated left to right,
 * unfused.  RenderMask reads a colour buffer nothing wrote (no fragment shader): undefined in the reference;
 * defined here as 255 where the depth test passed (z != 1), else 0. */
static float edge_linearize(uint32_t q, float zn, float zf) {
    if (q == ZMAX24) return -1.0f;                 /* if (z == 1) return float(-1) */
    float z = (float)((double)q / (double)ZMAX24); /* the depth texel as the sampler returns it */
    float a = 2.0f * zn;
    a = a * zf;
    float b = 2.0f * z;
    b = b - 1.0f;
    float c = zf - zn;
    b = b * c;
    float d = zf + zn;
    d = d - b;
    return a / d;
}

int vo_render_edge(const uint32_t *z24, int H, int W, float zn, float zf, uint8_t *out_edge) {
    const float lo = 0.05f, hi = 0.10f; /* kThreshLow / kThreshHigh */
    for (int j = 0; j < H; j++)
        for (int i = 0; i < W; i++) {
            uint8_t e = 0;
            /* pos < 5*d || pos > 1 - 5*d with pos = (i + 0.5)/W: columns 0..4 and W-5..W-1 (rows likewise) */
            if (i >= 5 && i < W - 5 && j >= 5 && j < H - 5 && z24[(size_t)j * W + i] != ZMAX24) {
                float v[9];
                int k = 0;
                for (int di = -1; di <= 1; di++)      /* value[0..8]: x outer, y inner, as in the shader */
                    for (int dj = -1; dj <= 1; dj++) v[k++] = edge_linearize(z24[(size_t)(j + dj) * W + (i + di)], zn, zf);
                float s = fabsf(v[1] - v[7]);
                s = s + fabsf(v[5] - v[3]);
                s = s + fabsf(v[0] - v[8]);
                s = s + fabsf(v[2] - v[6]);
                float delta = 0.25f * s;
                float c;
                if (delta < lo) c = 0.0f;
                else if (delta >= hi) c = 1.0f;
                else c = (delta - lo) / (hi - lo);
                e = (uint8_t)lrintf(c * 255.0f);
            }
            out_edge[(size_t)j * W + i] = e;
        }
    return 0;
}

int vo_render_mask(const uint32_t *z24, int H, int W, uint8_t *out_mask) {
    for (size_t p = 0; p < (size_t)H * W; p++) out_mask[p] = z24[p] != ZMAX24 ? 0 : 255;
    return 0;
}
