/*
 * oracle_force.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, fp64) of the force and energy loops of the reference
 * hot path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this; nothing under nbody-gnn-hpc_b200/ does.
 *
 * Follows, statement for statement:
 *   compute_accelerations_direct   /root/reference/src/hpc/nbody.py:22-66
 *   compute_total_energy           /root/reference/src/hpc/nbody.py:101-130
 *
 * This translation unit is compiled with -O3 -ffast-math -fopenmp so the
 * compiler is free to do what Numba's @jit(parallel=True, fastmath=True) lets
 * LLVM do to the reference (nbody.py:22): vectorise the j loop, contract
 * mul+add into FMA and reassociate the sums.  The *_strict variants below are
 * compiled in oracle_strict.c without those licences and give the source-order
 * semantics; the gap between the two is the reference's own rounding envelope.
 *
 * Parity pin: tests/test_oracle_golden.py checks these functions against
 * vectors produced by executing the reference's Numba functions
 * (tests/golden/make_golden.py).  The reference itself holds no tests or
 * golden vectors for this path.
 */
#include <math.h>
#include <stddef.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_G 6.67430e-11 /* nbody.py:18 */

/* masses may arrive as float64 or float32 (generate_data.py:109 hands float32);
 * Numba promotes G*masses[j] to float64 from the rounded value either way. */
static inline double mass_at(const void *masses, int masses_are_f32, int j)
{
    return masses_are_f32 ? (double)((const float *)masses)[j]
                          : ((const double *)masses)[j];
}

/* nbody.py:38-66.  i0/n_i select a slab of i rows (the whole system when
 * i0=0, n_i=n); rows outside the slab are not touched.  acc is (n,3). */
void oracle_accel_direct_rows(const double *pos, const void *masses,
                              int masses_are_f32, int n, double softening,
                              int i0, int n_i, double *acc)
{
    const double eps2 = softening * softening;
#pragma omp parallel for schedule(static)
    for (int i = i0; i < i0 + n_i; ++i) {               /* prange, nbody.py:41 */
        double ax = 0.0, ay = 0.0, az = 0.0;            /* :42 */
        const double xi = pos[3 * i + 0];               /* :43 */
        const double yi = pos[3 * i + 1];
        const double zi = pos[3 * i + 2];
        for (int j = 0; j < n; ++j) {                   /* :45 */
            if (i != j) {                               /* :46 */
                const double dx = pos[3 * j + 0] - xi;  /* :47 */
                const double dy = pos[3 * j + 1] - yi;  /* :48 */
                const double dz = pos[3 * j + 2] - zi;  /* :49 */
                const double r2 = dx * dx + dy * dy + dz * dz + eps2; /* :52 */
                const double r = sqrt(r2);              /* :53 */
                const double r3 = r * r2;               /* :54 */
                const double factor =
                    ORACLE_G * mass_at(masses, masses_are_f32, j) / r3; /* :57 */
                ax += factor * dx;                      /* :58 */
                ay += factor * dy;                      /* :59 */
                az += factor * dz;                      /* :60 */
            }
        }
        acc[3 * i + 0] = ax;                            /* :62-64 */
        acc[3 * i + 1] = ay;
        acc[3 * i + 2] = az;
    }
}

void oracle_accel_direct(const double *pos, const void *masses,
                         int masses_are_f32, int n, double softening,
                         double *acc)
{
    oracle_accel_direct_rows(pos, masses, masses_are_f32, n, softening, 0, n, acc);
}

/* Same loop without the OpenMP team: what one generate_data.py worker runs
 * (NUMBA_NUM_THREADS=1, generate_data.py:16-19).  Used by oracle_ensemble_run. */
void oracle_accel_direct_serial(const double *pos, const void *masses,
                                int masses_are_f32, int n, double softening,
                                double *acc)
{
    const double eps2 = softening * softening;
    for (int i = 0; i < n; ++i) {
        double ax = 0.0, ay = 0.0, az = 0.0;
        const double xi = pos[3 * i + 0], yi = pos[3 * i + 1], zi = pos[3 * i + 2];
        for (int j = 0; j < n; ++j) {
            if (i != j) {
                const double dx = pos[3 * j + 0] - xi;
                const double dy = pos[3 * j + 1] - yi;
                const double dz = pos[3 * j + 2] - zi;
                const double r2 = dx * dx + dy * dy + dz * dz + eps2;
                const double r = sqrt(r2);
                const double r3 = r * r2;
                const double factor =
                    ORACLE_G * mass_at(masses, masses_are_f32, j) / r3;
                ax += factor * dx;
                ay += factor * dy;
                az += factor * dz;
            }
        }
        acc[3 * i + 0] = ax;
        acc[3 * i + 1] = ay;
        acc[3 * i + 2] = az;
    }
}

/* nbody.py:101-130: single-threaded in the reference (no parallel=True);
 * out = {kinetic, potential, total}. */
void oracle_total_energy(const double *pos, const double *vel,
                         const void *masses, int masses_are_f32, int n,
                         double softening, double *out)
{
    const double eps2 = softening * softening;
    double kinetic = 0.0;                                /* :115 */
    for (int i = 0; i < n; ++i) {                        /* :116 */
        const double v2 = vel[3 * i] * vel[3 * i] + vel[3 * i + 1] * vel[3 * i + 1] +
                          vel[3 * i + 2] * vel[3 * i + 2]; /* :117 */
        kinetic += 0.5 * mass_at(masses, masses_are_f32, i) * v2; /* :118 */
    }
    double potential = 0.0;                              /* :121 */
    for (int i = 0; i < n; ++i) {                        /* :122 */
        const double mi = mass_at(masses, masses_are_f32, i);
        for (int j = i + 1; j < n; ++j) {                /* :123 */
            const double dx = pos[3 * j + 0] - pos[3 * i + 0]; /* :124 */
            const double dy = pos[3 * j + 1] - pos[3 * i + 1]; /* :125 */
            const double dz = pos[3 * j + 2] - pos[3 * i + 2]; /* :126 */
            const double r = sqrt(dx * dx + dy * dy + dz * dz + eps2); /* :127 */
            potential -= ORACLE_G * mi * mass_at(masses, masses_are_f32, j) / r; /* :128 */
        }
    }
    out[0] = kinetic;
    out[1] = potential;
    out[2] = kinetic + potential;                        /* :130 */
}

/* Parallel version of the same sums for the large-N energy-drift checks the
 * reference cannot finish; same pairs, per-row partials
 * reduced in row order. */
void oracle_total_energy_parallel(const double *pos, const double *vel,
                                  const void *masses, int masses_are_f32, int n,
                                  double softening, double *out)
{
    const double eps2 = softening * softening;
    double kinetic = 0.0, potential = 0.0;
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : kinetic, potential)
    for (int i = 0; i < n; ++i) {
        const double mi = mass_at(masses, masses_are_f32, i);
        const double v2 = vel[3 * i] * vel[3 * i] + vel[3 * i + 1] * vel[3 * i + 1] +
                          vel[3 * i + 2] * vel[3 * i + 2];
        kinetic += 0.5 * mi * v2;
        double row = 0.0;
        for (int j = i + 1; j < n; ++j) {
            const double dx = pos[3 * j + 0] - pos[3 * i + 0];
            const double dy = pos[3 * j + 1] - pos[3 * i + 1];
            const double dz = pos[3 * j + 2] - pos[3 * i + 2];
            const double r = sqrt(dx * dx + dy * dy + dz * dz + eps2);
            row += ORACLE_G * mi * mass_at(masses, masses_are_f32, j) / r;
        }
        potential -= row;
    }
    out[0] = kinetic;
    out[1] = potential;
    out[2] = kinetic + potential;
}

/* torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU baseline asks for the cores it really has. */
void oracle_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int oracle_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
