/*
 * TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's Lennard-Jones
 * MD step.  Nothing in the product path (lennard-jones-cuda_b200/, include/)
 * links, loads or calls this file; only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs do, and only as the checker
 * or the timed CPU baseline.
 *
 * Parity status: PINNED.  The reference publishes no golden vectors
 * (SURVEY.md §4), but its CPU path compiles here unmodified
 * (oracle/Makefile -> oracle/_ref/libljmd_ref.so).  tests/test_oracle.py checks
 * this restatement against that library bit-for-bit on seeded snapshots, and
 * against the fixtures under tests/golden/ that oracle/make_golden.py
 * generated from it.
 *
 * Build flags matter: -O2 -ffp-contract=off, no -march: the
 * float expression for r^2 must NOT be fused, or RDF bins change.
 *
 * Every function cites the reference lines it restates
 * (paths relative to /root/reference/src/library/).
 *
 * Array layout is the reference's: AoS float[4N], (x,y,z,w) per particle
 * (MDSystem.h:72-74, MDSystem.cpp:86-88).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define LJ_RDF_BINS 256

/* MDSystem.cpp:732-739 — round half away from zero through a float add. */
static int fast_round(float x)
{
  if (x > 0) return (int)(x + 0.5f);
  else       return (int)(x - 0.5f);
}

/* MDSystem.cpp:93-95 — RDF bin width in r^2 (0.1f for every N >= 400). */
float ljo_rdf_dr2(int N)
{
  float dr2 = (float)fmax(0.2 * sqrt(100. / N), 0.05);
  if (250 * dr2 < 25.0) dr2 = (float)(25.0 / 250);
  return dr2;
}

/* MDSystem.cpp:70 */
double ljo_box_length(int N, double rho) { return pow(N / rho, 1. / 3.); }

/*
 * MDSystem.cpp:252-310 — CPU all-pairs force / potential / virial / RDF.
 * bc: 0 periodic (minimum image), 1 hard wall, 2 none.  Only bc==0 images.
 * out: frc[4N] (xyz written, w untouched), scal[0]=V, scal[1]=P(virial part),
 * scal[2]=Pshear(configurational part), rdf[256] (int32).
 */
void ljo_forces(int N, const float* pos, double L, int bc, float rdf_dr2,
                float* frc, double* scal, int* rdf)
{
  double r2, r6;
  double V = 0., P = 0., Pshear = 0.;
  int i, j;
  memset(rdf, 0, LJ_RDF_BINS * sizeof(int));            /* :254 */
  for (i = 0; i < 4 * N; i += 4) {
    frc[i] = 0.f; frc[i + 1] = 0.f; frc[i + 2] = 0.f;   /* :262-264 */
    for (j = 0; j < 4 * N; j += 4) {
      if (j != i) {                                      /* :267 */
        float rx = pos[i] - pos[j];                      /* :269-271 */
        float ry = pos[i + 1] - pos[j + 1];
        float rz = pos[i + 2] - pos[j + 2];
        if (bc == 0) {                                   /* :273-277 */
          rx = (float)(rx - L * fast_round((float)(rx / L)));
          ry = (float)(ry - L * fast_round((float)(ry / L)));
          rz = (float)(rz - L * fast_round((float)(rz / L)));
        }
        r2 = rx * rx + ry * ry + rz * rz;                /* :279 float, widened */
        {
          int indrdf = (int)floor(r2 / rdf_dr2);         /* :282 */
          if ((size_t)indrdf < (size_t)LJ_RDF_BINS)      /* :283 int vs size_t */
            rdf[indrdf]++;
        }
        {
          double fijx, fijy, fijz;
          r2 = 1 / r2;                                   /* :289 */
          r6 = r2 * r2 * r2;                             /* :290 */
          fijx = r2 * (12 * r6 * r6 - 6 * r6) * rx;      /* :291-293 */
          fijy = r2 * (12 * r6 * r6 - 6 * r6) * ry;
          fijz = r2 * (12 * r6 * r6 - 6 * r6) * rz;
          frc[i]     = (float)(frc[i] + fijx);           /* :294-296 float accumulator */
          frc[i + 1] = (float)(frc[i + 1] + fijy);
          frc[i + 2] = (float)(frc[i + 2] + fijz);
          V += (1 * r6 * r6 - 1 * r6);                   /* :297 */
          P += rx * fijx + ry * fijy + rz * fijz;        /* :298 */
          Pshear += -rx * fijy;                          /* :299 */
        }
      }
    }
    frc[i] *= 4; frc[i + 1] *= 4; frc[i + 2] *= 4;       /* :303-305 */
  }
  P *= 4. / 3. / 2.;                                     /* :307 */
  V *= 4. / 2.;                                          /* :308 */
  Pshear *= 4. / 2.;                                     /* :309 */
  scal[0] = V; scal[1] = P; scal[2] = Pshear;
}

/*
 * FP64 arbiter: the same physics with every operation
 * and every accumulator in double, minimum image by rint().  Not a restatement
 * of reference arithmetic — used to judge which of two FP32 answers is closer
 * and to build the per-particle normalisation sum_j |f_ij| for the tolerance.
 * out: frc[3N] doubles (x,y,z, x4 applied), fabs_sum[2N]: [0..N) = 4*sum_j |f_ij| (net pair forces),
 * [N..2N) = 4*sum_j (12 r^-13 + 6 r^-7) (repulsive + attractive magnitudes: the terms any evaluation has to
 * cancel; near the potential minimum r = 2^(1/6) the net pair force vanishes while its two terms do not),
 * scal[0]=V, scal[1]=P(virial part) with the reference's prefactors;
 * scal[2]=sum of |pair potential| terms, scal[3]=sum of |pair virial| terms (same prefactors): the
 * scales against which "relative" errors of V and P are judged when V or P nearly cancels.
 */
void ljo_forces_f64(int N, const float* pos, double L, int bc,
                    double* frc, double* fabs_sum, double* scal)
{
  double V = 0., P = 0., Vabs = 0., Pabs = 0.;
  int i, j;
  for (i = 0; i < N; ++i) {
    double fx = 0., fy = 0., fz = 0., fa = 0., ft = 0.;
    const double xi = pos[4 * i], yi = pos[4 * i + 1], zi = pos[4 * i + 2];
    for (j = 0; j < N; ++j) {
      double rx, ry, rz, r2, ir2, r6, s;
      if (j == i) continue;
      rx = xi - pos[4 * j]; ry = yi - pos[4 * j + 1]; rz = zi - pos[4 * j + 2];
      if (bc == 0) {
        rx -= L * rint(rx / L); ry -= L * rint(ry / L); rz -= L * rint(rz / L);
      }
      r2 = rx * rx + ry * ry + rz * rz;
      ir2 = 1. / r2; r6 = ir2 * ir2 * ir2;
      s = ir2 * (12. * r6 * r6 - 6. * r6);
      fx += s * rx; fy += s * ry; fz += s * rz;
      fa += fabs(s) * sqrt(r2);
      ft += ir2 * (12. * r6 * r6 + 6. * r6) * sqrt(r2);
      V += r6 * r6 - r6;
      P += s * r2;
      Vabs += r6 * r6 + r6;
      Pabs += fabs(s * r2);
    }
    frc[3 * i] = 4. * fx; frc[3 * i + 1] = 4. * fy; frc[3 * i + 2] = 4. * fz;
    fabs_sum[i] = 4. * fa;
    fabs_sum[N + i] = 4. * ft;
  }
  scal[0] = V * 4. / 2.;
  scal[1] = P * 4. / 3. / 2.;
  scal[2] = Vabs * 4. / 2.;
  scal[3] = Pabs * 4. / 3. / 2.;
}

/*
 * FP64 arbiter on a SUBSAMPLE of the particles, for sizes where the O(N^2) arbiter above takes hours
 * (N = 65 536 ... 1 048 576): the force on each of the nsub particles idx[k] from all N, every operation and
 * accumulator in double, image rule as in ljo_forces_f64 (the reference's MDSystem.cpp:269-277 in exact
 * arithmetic).  O(nsub * N), split over `threads` host threads (independent particles; no shared writes).
 * out: frc[3*nsub] (x4 applied), fterm[nsub] = 4*sum_j (12 r^-13 + 6 r^-7) (the tolerance scale),
 * pe[2*nsub] = per particle {sum_j (r^-12 - r^-6) (what force.w carries, MDSystem.cu:52), sum_j (r^-12 + r^-6)}.
 */
#include <pthread.h>
typedef struct {
  int N, bc, k0, k1;
  const float* pos;
  const int* idx;
  double L;
  double *frc, *fterm, *pe;
} ljo_sub_job;

static void* ljo_sub_worker(void* arg)
{
  const ljo_sub_job* q = (const ljo_sub_job*)arg;
  int k, j;
  for (k = q->k0; k < q->k1; ++k) {
    const int i = q->idx[k];
    const double xi = q->pos[4 * i], yi = q->pos[4 * i + 1], zi = q->pos[4 * i + 2];
    double fx = 0., fy = 0., fz = 0., ft = 0., pe = 0., pa = 0.;
    for (j = 0; j < q->N; ++j) {
      double rx, ry, rz, r2, ir2, r6, s;
      if (j == i) continue;
      rx = xi - q->pos[4 * j]; ry = yi - q->pos[4 * j + 1]; rz = zi - q->pos[4 * j + 2];
      if (q->bc == 0) {
        rx -= q->L * rint(rx / q->L); ry -= q->L * rint(ry / q->L); rz -= q->L * rint(rz / q->L);
      }
      r2 = rx * rx + ry * ry + rz * rz;
      ir2 = 1. / r2; r6 = ir2 * ir2 * ir2;
      s = ir2 * (12. * r6 * r6 - 6. * r6);
      fx += s * rx; fy += s * ry; fz += s * rz;
      ft += ir2 * (12. * r6 * r6 + 6. * r6) * sqrt(r2);
      pe += r6 * r6 - r6;
      pa += r6 * r6 + r6;
    }
    q->frc[3 * k] = 4. * fx; q->frc[3 * k + 1] = 4. * fy; q->frc[3 * k + 2] = 4. * fz;
    q->fterm[k] = 4. * ft;
    q->pe[2 * k] = pe;
    q->pe[2 * k + 1] = pa;
  }
  return 0;
}

void ljo_forces_f64_subset(int N, const float* pos, double L, int bc, int nsub, const int* idx, int threads,
                           double* frc, double* fterm, double* pe)
{
  enum { MAXT = 64 };
  pthread_t th[MAXT];
  ljo_sub_job job[MAXT];
  int t, nt = threads < 1 ? 1 : (threads > MAXT ? MAXT : threads);
  if (nt > nsub) nt = nsub > 0 ? nsub : 1;
  for (t = 0; t < nt; ++t) {
    job[t].N = N; job[t].bc = bc; job[t].pos = pos; job[t].idx = idx; job[t].L = L;
    job[t].frc = frc; job[t].fterm = fterm; job[t].pe = pe;
    job[t].k0 = (int)((long long)nsub * t / nt);
    job[t].k1 = (int)((long long)nsub * (t + 1) / nt);
  }
  for (t = 1; t < nt; ++t) pthread_create(&th[t], 0, ljo_sub_worker, &job[t]);
  ljo_sub_worker(&job[0]);
  for (t = 1; t < nt; ++t) pthread_join(th[t], 0);
}

/* MDSystem.cpp:361-373 */
double ljo_kinetic_temperature(int N, const float* vel)
{
  double ret = 0.;
  int i;
  for (i = 0; i < 4 * N; i += 4)
    ret += (vel[i] * vel[i] + vel[i + 1] * vel[i + 1] + vel[i + 2] * vel[i + 2]);
  ret *= 1. / 3. / N;
  return ret;
}

/* MDSystem.cpp:406-436 */
void ljo_apply_boundary(int N, float* pos, float* vel, double L, int bc)
{
  int i, k;
  if (bc == 2) return;                                   /* :409-412 */
  if (bc == 0) {                                         /* :414-424 */
    for (i = 0; i < 4 * N; i += 4)
      for (k = 0; k < 3; ++k) {
        if (pos[i + k] < 0.) pos[i + k] = (float)(pos[i + k] + L);
        if (pos[i + k] > L)  pos[i + k] = (float)(pos[i + k] - L);
      }
  } else {                                               /* :425-435 */
    for (i = 0; i < 4 * N; i += 4)
      for (k = 0; k < 3; ++k) {
        if (pos[i + k] < 0. && vel[i + k] < 0) vel[i + k] = -vel[i + k];
        if (pos[i + k] > L  && vel[i + k] > 0) vel[i + k] = -vel[i + k];
      }
  }
}

/*
 * MDSystem.cpp:325-359, CPU branch.  In: vel, V, and P/Pshear as left by
 * ljo_forces.  out[0..5] = U, T, K, V, P, Pshear (finished values).
 */
void ljo_parameters(int N, double rho, const float* vel,
                    double V, double Pvirial, double Pshear_conf, double* out)
{
  double K = 0., T, P = Pvirial, U, Pshear = Pshear_conf;
  int i;
  for (i = 0; i < 4 * N; i += 4) {                       /* :332-336 */
    K += (vel[i] * vel[i] + vel[i + 1] * vel[i + 1] + vel[i + 2] * vel[i + 2]) / 2.;
    Pshear += -vel[i] * vel[i + 1];
  }
  T = 2. * K / 3. / N;                                   /* :348 */
  P += N * T;                                            /* :349 */
  P /= (N / rho);                                        /* :350 */
  U = K + V;                                             /* :351 */
  Pshear /= (N / rho);                                   /* :353 */
  out[0] = U; out[1] = T; out[2] = K; out[3] = V; out[4] = P; out[5] = Pshear;
}

/*
 * MDSystem.cpp:438-583 — one Integrate(dt).  State in/out: pos, vel, frc
 * (frc holds f(t) on entry, f(t+dt) on exit), scal[0..5] as ljo_parameters,
 * rdf[256].  canonical: 0 EVN (:442-464), 1 TVN (:465-510).
 */
void ljo_integrate(int N, double rho, double L, double T0, int canonical, int bc,
                   float rdf_dr2, double dt, float* pos, float* vel, float* frc,
                   double* scal, int* rdf)
{
  double fs[3];
  int i, k;
  if (!canonical) {
    for (i = 0; i < 4 * N; i += 4) {
      float f[3]; f[0] = frc[i]; f[1] = frc[i + 1]; f[2] = frc[i + 2];
      for (k = 0; k < 3; ++k) {
        pos[i + k] = (float)(pos[i + k] + (dt * vel[i + k] + dt * dt * f[k] / 2.));  /* :447-449 */
        vel[i + k] = (float)(vel[i + k] + dt * f[k] / 2.);                             /* :451-453 */
      }
    }
    ljo_forces(N, pos, L, bc, rdf_dr2, frc, fs, rdf);                                  /* :456 */
    for (i = 0; i < 4 * N; i += 4)
      for (k = 0; k < 3; ++k)
        vel[i + k] = (float)(vel[i + k] + dt * frc[i + k] / 2.);                       /* :460-462 */
  } else {
    float* tF = (float*)malloc(sizeof(float) * 4 * (size_t)N);
    float* tV = (float*)malloc(sizeof(float) * 4 * (size_t)N);
    double Tkin, chi;
    for (i = 0; i < 4 * N; i += 4) {
      float f[3]; f[0] = frc[i]; f[1] = frc[i + 1]; f[2] = frc[i + 2];
      for (k = 0; k < 3; ++k) {
        pos[i + k] = (float)(pos[i + k] + (dt * vel[i + k] + dt * dt * f[k] / 2.));  /* :473-475 */
        tF[i + k] = 0.5f * frc[i + k];                                                 /* :477-479 */
      }
    }
    ljo_forces(N, pos, L, bc, rdf_dr2, frc, fs, rdf);                                  /* :482 */
    for (i = 0; i < 4 * N; i += 4)
      for (k = 0; k < 3; ++k) {
        tF[i + k] += 0.5f * frc[i + k];                                                /* :486-488 */
        tV[i + k] = (float)(vel[i + k] + dt * tF[i + k] / 2.);                         /* :493-495 */
      }
    Tkin = ljo_kinetic_temperature(N, tV);                                             /* :498 */
    chi = sqrt(T0 / Tkin);                                                             /* :499 */
    for (i = 0; i < 4 * N; i += 4)
      for (k = 0; k < 3; ++k)
        vel[i + k] = (float)((2. * chi - 1.) * vel[i + k] + chi * dt * tF[i + k]);     /* :503-505 */
    free(tF); free(tV);
  }
  ljo_apply_boundary(N, pos, vel, L, bc);                                              /* :578 */
  ljo_parameters(N, rho, vel, fs[0], fs[1], fs[2], scal);                              /* :579 */
}

/*
 * MDSystem.cpp:651-669 / :676-688 — speed histogram counts.
 * bin = (int)(sqrt((double)(float)(vx^2+vy^2+vz^2)) / step), dropped if >= nbins.
 */
void ljo_velocity_histogram(int N, const float* vel, double step, int nbins, int* dens)
{
  int i;
  memset(dens, 0, sizeof(int) * (size_t)nbins);
  for (i = 0; i < 4 * N; i += 4) {
    int tind = (int)(sqrt(vel[i] * vel[i] + vel[i + 1] * vel[i + 1] + vel[i + 2] * vel[i + 2]) / step);
    if (tind < nbins) dens[tind]++;
  }
}

/* MDSystem.cpp:633-649 — g(r) points from the r^2 histogram. */
void ljo_rdf_curve(int N, double L, float rdf_dr2, const int* rdf, double* r, double* g)
{
  const double PI = 3.141592653589793238462643;
  double n0 = N / L / L / L;
  int ir;
  for (ir = 0; ir < LJ_RDF_BINS; ++ir) {
    double r2 = (ir + 0.5) * rdf_dr2;
    double rr = sqrt(r2);
    r[ir] = rr;
    g[ir] = rdf[ir] / rdf_dr2 / 2. / PI / rr / n0 / (double)N;
  }
}

/*
 * MDSystem.cpp:147-168 — the simple-cubic start lattice (positions only; the
 * reference's velocities are time-seeded and are never reproduced — snapshots
 * carry their own seeded velocities).
 */
void ljo_lattice(int N, double L, float* pos)
{
  int Nsingle = (int)ceil(pow(N, 1. / 3.));
  double dL = L / Nsingle;
  int iN;
  for (iN = 0; iN < N; ++iN) {
    int ix = iN % Nsingle, iy = (iN / Nsingle) % Nsingle, iz = iN / (Nsingle * Nsingle);
    pos[4 * iN]     = (float)((ix + 0.5) * dL);
    pos[4 * iN + 1] = (float)((iy + 0.5) * dL);
    pos[4 * iN + 2] = (float)((iz + 0.5) * dL);
    pos[4 * iN + 3] = (float)(L / 150.f);
  }
}
