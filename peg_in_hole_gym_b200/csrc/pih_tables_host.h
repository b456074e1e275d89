// pih_tables_host.h - host side of the model tables: every translation unit that uses c_arm owns a copy of it
// (no relocatable device code) and uploads it with this function.  Internal.
#pragma once
#include <cstring>

#define FILL_ARM(T, P)                                                                          \
  do {                                                                                          \
    (T).nb = PIH_##P##_NB; (T).narm = PIH_##P##_NARM; (T).ee_body = PIH_##P##_EE_BODY;            \
    (T).nsph = PIH_##P##_NSPH; (T).nbox = PIH_##P##_NBOX;                                       \
    for (int i = 0; i < PIH_##P##_NB; ++i) {                                                    \
      (T).parent[i] = PIH_##P##_PARENT[i]; (T).jtype[i] = PIH_##P##_JTYPE[i];                    \
      unsigned a = 1u << i;                                                                     \
      for (int p = PIH_##P##_PARENT[i]; p >= 0; p = PIH_##P##_PARENT[p]) a |= 1u << p;           \
      (T).anc[i] = a;                                                                           \
      for (int r = 0; r < 3; ++r) {                                                             \
        for (int c = 0; c < 3; ++c) {                                                           \
          (T).R0[i][r][c] = (float)PIH_##P##_R0[i][3 * r + c];                                   \
          (T).Irot[i][r][c] = (float)PIH_##P##_IROT[i][3 * r + c];                               \
        }                                                                                       \
        (T).t0[i][r] = (float)PIH_##P##_T0[i][r]; (T).h[i][r] = (float)PIH_##P##_H[i][r];         \
      }                                                                                         \
      (T).mass[i] = (float)PIH_##P##_MASS[i]; (T).lo[i] = (float)PIH_##P##_LO[i]; (T).hi[i] = (float)PIH_##P##_HI[i]; \
      (T).damp[i] = (float)PIH_##P##_DAMP[i]; (T).maxf[i] = (float)PIH_##P##_MAXF[i];            \
      (T).kp[i] = (float)PIH_##P##_KP[i]; (T).rest[i] = (float)PIH_##P##_REST[i];                \
    }                                                                                           \
    for (int r = 0; r < 3; ++r) {                                                               \
      for (int c = 0; c < 3; ++c) (T).ee_R[r][c] = (float)PIH_##P##_EE_R[3 * r + c];             \
      (T).ee_t[r] = (float)PIH_##P##_EE_T[r];                                                   \
    }                                                                                           \
    for (int s = 0; s < PIH_MAX_SPH; ++s) {                                                     \
      (T).sph_body[s] = PIH_##P##_SPH_BODY[s]; (T).sph_r[s] = (float)PIH_##P##_SPH_R[s];         \
      for (int r = 0; r < 3; ++r) (T).sph_c[s][r] = (float)PIH_##P##_SPH_C[s][r];                \
    }                                                                                           \
    for (int s = 0; s < PIH_MAX_BOX; ++s) {                                                     \
      (T).box_body[s] = PIH_##P##_BOX_BODY[s];                                                  \
      for (int r = 0; r < 3; ++r) {                                                             \
        (T).box_c[s][r] = (float)PIH_##P##_BOX_C[s][r]; (T).box_h[s][r] = (float)PIH_##P##_BOX_H[s][r]; \
        for (int c = 0; c < 3; ++c) (T).box_R[s][r][c] = (float)PIH_##P##_BOX_R[s][3 * r + c];   \
      }                                                                                         \
    }                                                                                           \
  } while (0)

static cudaError_t upload_tables() {
  static ArmTables host[2];
  memset(host, 0, sizeof(host));
  FILL_ARM(host[PIH_ARM_PANDA], PANDA);
  FILL_ARM(host[PIH_ARM_UR5], UR5);
  return cudaMemcpyToSymbol(c_arm, host, sizeof(host));
}

