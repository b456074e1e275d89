// pih_render.cuh - batched camera of PegInHole.render (envs/peg_in_hole.py:276-304): a ray caster over the analytic
// primitives of the scene, one pixel per thread, one float4 (depth, r, g, b) store per pixel.  HBM-write bound
// (16 B per pixel, 1.44 MB per 300x300 env image).  Same formulas as the oracle's render_pixel.
#pragma once
#include "pih_step.cuh"

namespace pih {

#define PIH_RENDER_PPT 8

// Conservative screen rectangle of an oriented box (centre c, axes = columns of R, half extents h): the box is clipped
// against the plane just in front of the eye (no ray parameter below the near distance is accepted, so nothing behind
// it can be hit), and the rectangle bounds the projections of the corners in front of that plane and of the points
// where edges cross it.  Unlike the bounding-sphere bound this stays tight for bodies next to the camera (the fingers
// and a grasped peg sit within centimetres of the eye, mostly above it).  Never culls a hit.
PIH_D void screen_rect_box(v3 e, const m3& R, v3 c, v3 h, int W, int H, float* out4) {
  const float ith = 1.0f / (float)PIH_CAM_TAN_HALF_FOV, zc = 0.5f * (float)PIH_CAM_NEAR;
  h = mk(h.x * 1.02f + 1e-4f, h.y * 1.02f + 1e-4f, h.z * 1.02f + 1e-4f);
  float px[8], py[8], pz[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    v3 l = mk((k & 1) ? h.x : -h.x, (k & 2) ? h.y : -h.y, (k & 4) ? h.z : -h.z);
    v3 w = mv_add(R, l, c);
    px[k] = w.x - e.x; py[k] = w.y - e.y; pz[k] = e.z - w.z;   // pz: depth along the viewing direction
  }
  float xlo = 1e30f, xhi = -1e30f, ylo = 1e30f, yhi = -1e30f;
  auto add = [&](float x, float y, float z) {
    float iz = ith / z;
    float u = x * iz, v = y * iz;
    xlo = fminf(xlo, u); xhi = fmaxf(xhi, u); ylo = fminf(ylo, v); yhi = fmaxf(yhi, v);
  };
#pragma unroll
  for (int k = 0; k < 8; ++k)
    if (pz[k] > zc) add(px[k], py[k], pz[k]);
#pragma unroll
  for (int k = 0; k < 8; ++k)
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const int m = k ^ (1 << a);
      if (m < k) continue;                                      // each of the 12 edges once
      if ((pz[k] > zc) != (pz[m] > zc)) {
        float t = (zc - pz[k]) / (pz[m] - pz[k]);
        add(fmaf(t, px[m] - px[k], px[k]), fmaf(t, py[m] - py[k], py[k]), zc);
      }
    }
  float r0 = 1.0f, r1 = 0.0f, c0 = 1.0f, c1 = 0.0f;              // empty rectangle: the whole box is behind the eye
  if (xlo <= xhi) {
    // clamp in normalised coordinates first (clipped points can project arbitrarily far out)
    xlo = fmaxf(xlo, -2.0f); xhi = fminf(xhi, 2.0f); ylo = fmaxf(ylo, -2.0f); yhi = fminf(yhi, 2.0f);
    c0 = fmaxf(0.0f, floorf((xlo + 1.0f) * 0.5f * (float)W - 0.5f) - 1.0f);
    c1 = fminf((float)(W - 1), ceilf((xhi + 1.0f) * 0.5f * (float)W - 0.5f) + 1.0f);
    r0 = fmaxf(0.0f, floorf((1.0f - yhi) * 0.5f * (float)H - 0.5f) - 1.0f);
    r1 = fminf((float)(H - 1), ceilf((1.0f - ylo) * 0.5f * (float)H - 0.5f) + 1.0f);
  }
  out4[0] = r0; out4[1] = r1; out4[2] = c0; out4[3] = c1;
}

// depth-buffer value and shaded colour of a hit (obj < 0: nothing hit) - the tail of the oracle's render_pixel
PIH_D float4 shade_pixel(int obj, float best, v3 nb) {
  const float fn = (float)PIH_CAM_FAR, nn = (float)PIH_CAM_NEAR;
  float depth = 1.0f, r = 255.0f, g = 255.0f, bl = 255.0f, shade = 1.0f;
  if (obj >= 0) {
    depth = fmaf(0.5f, (fn + nn) / (fn - nn) - (2.0f * fn * nn) / ((fn - nn) * best), 0.5f);
    r = (obj == 0) ? 153.0f : (obj == 3) ? 204.0f : 232.0f;
    g = (obj == 0) ? 128.0f : (obj == 3) ? 204.0f : 232.0f;
    bl = (obj == 0) ? 102.0f : (obj == 3) ? 204.0f : 232.0f;
    const v3 l = mk((float)0.2672612419124244, (float)0.1781741612749496, (float)0.9449111825230680);
    float nl = vdot(nb, l);
    nl = (nl < 0.0f) ? 0.0f : nl;
    shade = fmaf(0.65f, nl, 0.35f);
  }
  float4 px;
  px.x = depth;
  px.y = (float)(int)(r * shade + 0.5f);
  px.z = (float)(int)(g * shade + 0.5f);
  px.w = (float)(int)(bl * shade + 0.5f);
  return px;
}

// scene block per env: eye(3) Rfb(9) pfb(3) Rbox0(9) cbox0(3) Rbox1(9) cbox1(3) | 39: screen rects (r0,r1,c0,c1) of the
// free body, the hole and the two finger boxes
// latched != 0 (scripted task): the scene captured when the state machine entered state 2 (state slots 12..15) instead of
// the current one - the reference's `grasp_img` (peg_in_hole.py:65-66); envs that have not captured yet show the present
template <int ARM> __global__ void __launch_bounds__(128) k_scene(SimParams P, float* __restrict__ scene, int W, int H, int latched) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= P.n) return;
  constexpr int nb = ArmDims<ARM>::NB;
  const ArmTables& T = c_arm[ARM];
  EnvState S;
  load_state(P, e, S);
  if (latched) {
    float w[16];
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const float4 v = __ldg(&P.state[(size_t)(12 + s) * P.n + e]);
      w[4 * s] = v.x; w[4 * s + 1] = v.y; w[4 * s + 2] = v.z; w[4 * s + 3] = v.w;
    }
    if (w[12] != 0.0f || w[13] != 0.0f || w[14] != 0.0f || w[15] != 0.0f) {
#pragma unroll
      for (int k = 0; k < 9; ++k) S.q[k] = w[k];
#pragma unroll
      for (int k = 0; k < 3; ++k) S.ppos[k] = w[9 + k];
#pragma unroll
      for (int k = 0; k < 4; ++k) S.pquat[k] = w[12 + k];
    }
  }
  FK<ARM> F;
  fk<ARM, nb>(S.q, F);
  v3 eye; m3 Ree;
  ee_pose<ARM>(F, eye, Ree);
  float* o = scene + (size_t)PIH_SCENE_WORDS * e;
  o[0] = eye.x; o[1] = eye.y; o[2] = eye.z;
  m3 Rfb = quat_to_m3(S.pquat);
#pragma unroll
  for (int i = 0; i < 9; ++i) o[3 + i] = Rfb.m[i / 3][i % 3];
  o[12] = S.ppos[0]; o[13] = S.ppos[1]; o[14] = S.ppos[2];
  sfor<0, PIH_MAX_BOX>([&](auto b_) {
    constexpr int b = decltype(b_)::value;
    constexpr int body = (b < ArmGeom<ARM>::NBOX) ? box_body_of<ARM>(b) : 0;
    m3 Rb = mmul(F.Rw[body], ldm(T.box_R[b]));
    v3 cb = mv_add(F.Rw[body], ld3(T.box_c[b]), F.pw[body]);
#pragma unroll
    for (int i = 0; i < 9; ++i) o[15 + 12 * b + i] = Rb.m[i / 3][i % 3];
    o[24 + 12 * b] = cb.x; o[25 + 12 * b] = cb.y; o[26 + 12 * b] = cb.z;
    screen_rect_box(eye, Rb, cb, ld3(T.box_h[b]), W, H, o + 47 + 4 * b);
  });
  // free body: its own box, or the box around the peg cylinder (fb_half = (r, half length, r), axis = local y)
  screen_rect_box(eye, Rfb, mk(S.ppos[0], S.ppos[1], S.ppos[2]), mk(P.fb_half[0], P.fb_half[1], P.fb_half[2]), W, H, o + 39);
  {  // hole: box around the tube (axis = env x)
    m3 I3;
#pragma unroll
    for (int i = 0; i < 9; ++i) I3.m[i / 3][i % 3] = (i / 3 == i % 3) ? 1.0f : 0.0f;
    screen_rect_box(eye, I3, mk((float)PIH_HOLE_X, (float)PIH_HOLE_Y, (float)PIH_HOLE_Z),
                    mk((float)PIH_HOLE_HL, (float)PIH_HOLE_ROUT, (float)PIH_HOLE_ROUT), W, H, o + 43);
  }
  // words 55..58: the pixel of a ray that hits nothing but the table.  The camera looks straight down, so the ray
  // parameter of the table plane (eye-space depth) and its shading are the same for every pixel of the env.
  float4 bg = shade_pixel(eye.z - (float)PIH_TABLE_Z > (float)PIH_CAM_NEAR ? 0 : -1, eye.z - (float)PIH_TABLE_Z, mk(0.0f, 0.0f, 1.0f));
  o[55] = bg.x; o[56] = bg.y; o[57] = bg.z; o[58] = bg.w;
}

PIH_D bool ray_box(v3 o, v3 d, v3 h, float smin, float& s_hit, v3& n) {
  float oo[3] = {o.x, o.y, o.z}, dd[3] = {d.x, d.y, d.z}, hh[3] = {h.x, h.y, h.z};
  float s0 = smin, s1 = 1e30f;
  int ax0 = -1;
  float sg0 = 0.0f;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    if (fabsf(dd[a]) < 1e-12f) {
      if (fabsf(oo[a]) > hh[a]) return false;
      continue;
    }
    float inv = 1.0f / dd[a];
    float ta = (-hh[a] - oo[a]) * inv, tb = (hh[a] - oo[a]) * inv;
    float sgn = -1.0f;
    if (ta > tb) { float t = ta; ta = tb; tb = t; sgn = 1.0f; }
    if (ta > s0) { s0 = ta; ax0 = a; sg0 = sgn; }
    if (tb < s1) s1 = tb;
  }
  if (s0 > s1 || ax0 < 0) return false;
  s_hit = s0;
  n = mk(ax0 == 0 ? sg0 : 0.0f, ax0 == 1 ? sg0 : 0.0f, ax0 == 2 ? sg0 : 0.0f);
  return true;
}

PIH_D bool ray_cyl_y(v3 o, v3 d, float r, float hl, float smin, float& s_hit, v3& n) {
  bool hit = false;
  float best = 1e30f;
  float a = fmaf(d.z, d.z, d.x * d.x), b = fmaf(o.z, d.z, o.x * d.x), c = fmaf(o.z, o.z, o.x * o.x) - r * r;
  if (a > 1e-12f) {
    float disc = fmaf(b, b, -(a * c));
    if (disc >= 0.0f) {
      float s = (-b - sqrtf(disc)) / a;
      float y = fmaf(s, d.y, o.y);
      if (s > smin && fabsf(y) <= hl) { best = s; hit = true; float ir = 1.0f / r; n = mk(fmaf(s, d.x, o.x) * ir, 0.0f, fmaf(s, d.z, o.z) * ir); }
    }
  }
  if (fabsf(d.y) > 1e-12f) {
    float inv = 1.0f / d.y;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      float yc = k ? -hl : hl;
      float s = (yc - o.y) * inv;
      float x = fmaf(s, d.x, o.x), z = fmaf(s, d.z, o.z);
      if (s > smin && s < best && fmaf(z, z, x * x) <= r * r) { best = s; hit = true; n = mk(0.0f, k ? -1.0f : 1.0f, 0.0f); }
    }
  }
  s_hit = best;
  return hit;
}

PIH_D bool ray_tube_x(v3 o, v3 d, float smin, float& s_hit, v3& n) {
  const float rin = (float)PIH_HOLE_RIN, rout = (float)PIH_HOLE_ROUT, hl = (float)PIH_HOLE_HL;
  float a = fmaf(d.z, d.z, d.y * d.y), b = fmaf(o.z, d.z, o.y * d.y), q2 = fmaf(o.z, o.z, o.y * o.y);
  float sa = smin, sb = 1e30f;
  bool slab_first = false;
  float slab_sign = 0.0f;
  if (fabsf(d.x) > 1e-12f) {
    float inv = 1.0f / d.x;
    float ta = (-hl - o.x) * inv, tb = (hl - o.x) * inv, sg = -1.0f;
    if (ta > tb) { float t = ta; ta = tb; tb = t; sg = 1.0f; }
    if (ta > sa) { sa = ta; slab_first = true; slab_sign = sg; }
    sb = tb;
  } else if (fabsf(o.x) > hl) return false;
  if (!(a > 1e-12f)) return false;
  float disco = fmaf(b, b, -(a * (q2 - rout * rout)));
  if (disco < 0.0f) return false;
  float sq = sqrtf(disco), ia = 1.0f / a;
  float so1 = (-b - sq) * ia, so2 = (-b + sq) * ia;
  bool first_is_slab = slab_first;
  float s0 = sa, s1 = (sb < so2) ? sb : so2;
  if (so1 > s0) { s0 = so1; first_is_slab = false; }
  if (s0 > s1) return false;
  float disci = fmaf(b, b, -(a * (q2 - rin * rin)));
  float si1 = 1e30f, si2 = -1e30f;
  if (disci > 0.0f) { float sqi = sqrtf(disci); si1 = (-b - sqi) * ia; si2 = (-b + sqi) * ia; }
  if (s0 < si1 || s0 >= si2) {
    if (first_is_slab && s0 == smin) return false;
    s_hit = s0;
    if (first_is_slab) n = mk(slab_sign, 0.0f, 0.0f);
    else { float ir = 1.0f / rout; n = mk(0.0f, fmaf(s0, d.y, o.y) * ir, fmaf(s0, d.z, o.z) * ir); }
    return s0 > smin;
  }
  if (si2 < s1 && si2 > smin) {
    float ir = 1.0f / rin;
    s_hit = si2;
    n = mk(0.0f, -(fmaf(si2, d.y, o.y) * ir), -(fmaf(si2, d.z, o.z) * ir));
    return true;
  }
  return false;
}

// conservative sphere cull: can the ray o + s d come within `rad` of centre c?  (never rejects a true hit)
PIH_D bool near_sphere(v3 o, v3 d, v3 c, float rad) {
  v3 oc = vsub(c, o);
  float dd = vdot(d, d), t = vdot(oc, d);
  float perp2 = vdot(oc, oc) * dd - t * t;         // |oc x d|^2
  return perp2 <= rad * rad * dd * 1.1f + 1e-9f;
}

// One float4 (depth, r, g, b) per pixel, PIH_RENDER_PPT pixels per thread strided by the block size so that every warp
// store is 512 contiguous bytes.  Most pixels see only the table: a warp none of whose 32 pixels lies inside the screen
// rectangle of an object stores the env's background pixel (computed once by k_scene) without casting a ray, which is
// what lets the kernel run at the HBM write rate; the others run the full ray caster (same formulas as the oracle).
template <int ARM, int TASK> __global__ void __launch_bounds__(256) k_render(SimParams P, const float* __restrict__ scene, int W, int H,
                                                                       float4* __restrict__ out) {
  __shared__ float sc[PIH_SCENE_WORDS];
  const int e = blockIdx.y;
  if (threadIdx.x < PIH_SCENE_WORDS) sc[threadIdx.x] = scene[(size_t)PIH_SCENE_WORDS * e + threadIdx.x];
  __syncthreads();
  const ArmTables& T = c_arm[ARM];
  const float4 bg = make_float4(sc[55], sc[56], sc[57], sc[58]);
  const float invW = 1.0f / (float)W;
  const int npix = W * H;
  float4* __restrict__ img = out + (size_t)e * npix;
#define IN_RECT(base) (frow >= sc[(base)] && frow <= sc[(base) + 1] && fcol >= sc[(base) + 2] && fcol <= sc[(base) + 3])
#pragma unroll 1
  for (int it = 0; it < PIH_RENDER_PPT; ++it) {
    const int pix = (blockIdx.x * PIH_RENDER_PPT + it) * blockDim.x + threadIdx.x;
    const bool valid = pix < npix;
    // row = pix / W without an integer division: exact for images up to 4096 x 4096 (error 4e-5 against a margin 0.5/W)
    const int row = (int)(((float)pix + 0.5f) * invW), col = pix - row * W;
    const float frow = (float)row, fcol = (float)col;
    bool near_obj = valid && (IN_RECT(39) || (ArmGeom<ARM>::NBOX > 0 && IN_RECT(47)) || (ArmGeom<ARM>::NBOX > 1 && IN_RECT(51)));
    if (TASK != PIH_TASK_RANDOM_FLY) near_obj = near_obj || (valid && IN_RECT(43));
    float4 px = bg;
    if (__any_sync(0xffffffffu, near_obj)) {
      if (near_obj) {
        const float th = (float)PIH_CAM_TAN_HALF_FOV;
        float ndx = ((float)col + 0.5f) / (float)W * 2.0f - 1.0f;
        float ndy = 1.0f - ((float)row + 0.5f) / (float)H * 2.0f;
        v3 d = mk(ndx * th, ndy * th, -1.0f);
        v3 o = mk(sc[0], sc[1], sc[2]);
        const float smin = (float)PIH_CAM_NEAR;
        float best = 1e30f;
        v3 nb = mk(0.0f, 0.0f, 1.0f);
        int obj = -1;
        { float s = o.z - (float)PIH_TABLE_Z; if (s > smin) { best = s; obj = 0; } }
        if (IN_RECT(39)) {
          m3 Rfb;
#pragma unroll
          for (int i = 0; i < 9; ++i) Rfb.m[i / 3][i % 3] = sc[3 + i];
          v3 pfb = mk(sc[12], sc[13], sc[14]);
          v3 ol = mTv(Rfb, vsub(o, pfb)), dl = mTv(Rfb, d);
          float s; v3 n; bool h;
          if (TASK != PIH_TASK_RANDOM_FLY) h = ray_cyl_y(ol, dl, P.fb_half[0], P.fb_half[1], smin, s, n);
          else h = ray_box(ol, dl, mk(P.fb_half[0], P.fb_half[1], P.fb_half[2]), smin, s, n);
          if (h && s < best) { best = s; obj = 1; nb = mv(Rfb, n); }
        }
        if (TASK != PIH_TASK_RANDOM_FLY) {
          if (IN_RECT(43)) {
            v3 oh = mk(o.x - (float)PIH_HOLE_X, o.y - (float)PIH_HOLE_Y, o.z - (float)PIH_HOLE_Z);
            float s; v3 n;
            if (ray_tube_x(oh, d, smin, s, n) && s < best) { best = s; obj = 2; nb = n; }
          }
        }
#pragma unroll
        for (int b = 0; b < ArmGeom<ARM>::NBOX; ++b) {
          if (IN_RECT(47 + 4 * b)) {
            m3 Rb;
#pragma unroll
            for (int i = 0; i < 9; ++i) Rb.m[i / 3][i % 3] = sc[15 + 12 * b + i];
            v3 cb = mk(sc[24 + 12 * b], sc[25 + 12 * b], sc[26 + 12 * b]);
            v3 hb = ld3(T.box_h[b]);
            v3 ol = mTv(Rb, vsub(o, cb)), dl = mTv(Rb, d);
            float s; v3 n;
            if (ray_box(ol, dl, hb, smin, s, n) && s < best) { best = s; obj = 3; nb = mv(Rb, n); }
          }
        }
        px = shade_pixel(obj, best, nb);
      }
    }
    if (valid) __stcs(&img[pix], px);   // written once, never re-read by this kernel: streaming store
  }
#undef IN_RECT
}

// ---------------------------------------------------------------- grasp label images of random_grasp
// peg_in_hole.py:41-45,65-98: on entering FSM state 2 the reference fills a rotated rectangle (half extents 0.1 x 0.2 in
// normalised image units, angle = atan2 of the rotated random_vector) into pos_img (50), ang_img and wid_img with
// skimage.draw.polygon and returns [pos_img, sin(2 ang_img), cos(2 ang_img), wid_img] plus [x, y, angle in degrees,
// width, length].  Vertices and inside test (closed quad, pixel = integer point) are shared with the oracle.
struct LabelRect { float vx[4], vy[4], s2, c2, wid, len; bool any; };
PIH_D LabelRect label_rect(float ang, int W, int H) {
  LabelRect R;
  R.any = ang > (float)(0.5 * PIH_LABEL_NONE);
  float sa, ca;
  det_sincos(R.any ? ang : 0.0f, sa, ca);
  det_sincos(R.any ? 2.0f * ang : 0.0f, R.s2, R.c2);
  const float L = (float)PIH_LABEL_LENGTH, Wd = (float)PIH_LABEL_WIDTH, S0 = (float)W, S1 = (float)H;
  // a, c, b, d of peg_in_hole.py:78-83; first coordinate <-> image column, second <-> image row (pos_img[cc, rr])
  R.vx[0] = (1.0f + L * ca + Wd * sa) / 2.0f * S0; R.vy[0] = (1.0f - L * sa + Wd * ca) / 2.0f * S1;
  R.vx[1] = (1.0f - L * ca + Wd * sa) / 2.0f * S0; R.vy[1] = (1.0f + L * sa + Wd * ca) / 2.0f * S1;
  R.vx[2] = (1.0f - L * ca - Wd * sa) / 2.0f * S0; R.vy[2] = (1.0f + L * sa - Wd * ca) / 2.0f * S1;
  R.vx[3] = (1.0f + L * ca - Wd * sa) / 2.0f * S0; R.vy[3] = (1.0f - L * sa - Wd * ca) / 2.0f * S1;
  float dx = R.vx[0] - R.vx[3], dy = R.vy[0] - R.vy[3];
  R.wid = sqrtf(fmaf(dx, dx, dy * dy));
  dx = R.vx[0] - R.vx[1]; dy = R.vy[0] - R.vy[1];
  R.len = sqrtf(fmaf(dx, dx, dy * dy));
  return R;
}
PIH_D bool label_inside(const LabelRect& R, float x, float y) {
  bool pos = true, neg = true;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int m = (k + 1) & 3;
    float cr = fmaf(R.vx[m] - R.vx[k], y - R.vy[k], -((R.vy[m] - R.vy[k]) * (x - R.vx[k])));
    pos = pos && (cr >= 0.0f); neg = neg && (cr <= 0.0f);
  }
  return R.any && (pos || neg);
}
// labels float[n][4][H][W] (pos, sin, cos, wid); params float[n][5] (x, y, angle in degrees, width, length in pixels)
__global__ void __launch_bounds__(256) k_labels(SimParams P, int W, int H, float* __restrict__ labels, float* __restrict__ params) {
  const int e = blockIdx.y;
  const float ang = reinterpret_cast<const float*>(&P.state[(size_t)(PIH_S_GRASP_ANG >> 2) * P.n + e])[PIH_S_GRASP_ANG & 3];
  const LabelRect R = label_rect(ang, W, H);
  if (blockIdx.x == 0 && threadIdx.x == 0 && params) {
    float* o = params + 5 * (size_t)e;
    o[0] = 0.0f; o[1] = 0.0f;
    o[2] = R.any ? ang / (float)3.141592653589793 * 180.0f : 0.0f;
    o[3] = R.any ? R.wid : 0.0f; o[4] = R.any ? R.len : 0.0f;
  }
  const int npix = W * H;
  const float invW = 1.0f / (float)W;
  float* img = labels + (size_t)e * 4 * npix;
#pragma unroll 1
  for (int it = 0; it < PIH_RENDER_PPT; ++it) {
    const int pix = (blockIdx.x * PIH_RENDER_PPT + it) * blockDim.x + threadIdx.x;
    if (pix >= npix) return;
    const int row = (int)(((float)pix + 0.5f) * invW), col = pix - row * W;
    const bool in = label_inside(R, (float)col, (float)row);
    __stcs(&img[pix], in ? (float)PIH_LABEL_POS_VALUE : 0.0f);
    __stcs(&img[npix + pix], in ? R.s2 : 0.0f);
    __stcs(&img[2 * npix + pix], in ? R.c2 : 1.0f);
    __stcs(&img[3 * npix + pix], in ? R.wid : 0.0f);
  }
}

}  // namespace pih
