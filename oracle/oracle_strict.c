/*
 * oracle_strict.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * The parts of the reference hot path whose rounding is fixed by the source:
 * the NumPy leapfrog of NBodySimulator.step / run, and source-order variants of
 * the force loop.  Compiled with -O2 -fno-fast-math -ffp-contract=off so every
 * multiply and add below is rounded separately, exactly as NumPy's elementwise
 * ops are.
 *
 * Follows:
 *   NBodySimulator.step   /root/reference/src/hpc/nbody.py:202-218
 *   NBodySimulator.run    /root/reference/src/hpc/nbody.py:220-248
 *   get_state             /root/reference/src/hpc/nbody.py:250-259
 *   (ensemble driver)     /root/reference/scripts/generate_data.py:32-58,142-149
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this.
 */
#include <math.h>
#include <stddef.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_G 6.67430e-11 /* nbody.py:18 */

/* from oracle_force.c (fast-math build, like the reference's Numba kernel) */
void oracle_accel_direct(const double *pos, const void *masses, int masses_are_f32,
                         int n, double softening, double *acc);
void oracle_accel_direct_serial(const double *pos, const void *masses,
                                int masses_are_f32, int n, double softening,
                                double *acc);

static inline double mass_at(const void *masses, int masses_are_f32, int j)
{
    return masses_are_f32 ? (double)((const float *)masses)[j]
                          : ((const double *)masses)[j];
}

/* nbody.py:41-64 evaluated literally: ascending j (reverse=0) or descending j
 * (reverse=1), every operation rounded once, no contraction.  The difference
 * between the two orders is the summation-order noise floor. */
void oracle_accel_direct_strict(const double *pos, const void *masses,
                                int masses_are_f32, int n, double softening,
                                int reverse, double *acc)
{
    const double eps2 = softening * softening;
    for (int i = 0; i < n; ++i) {
        double ax = 0.0, ay = 0.0, az = 0.0;
        const double xi = pos[3 * i + 0], yi = pos[3 * i + 1], zi = pos[3 * i + 2];
        for (int jj = 0; jj < n; ++jj) {
            const int j = reverse ? (n - 1 - jj) : jj;
            if (i != j) {
                const double dx = pos[3 * j + 0] - xi;
                const double dy = pos[3 * j + 1] - yi;
                const double dz = pos[3 * j + 2] - zi;
                const double r2 = dx * dx + dy * dy + dz * dz + eps2;
                const double r = sqrt(r2);
                const double r3 = r * r2;
                const double factor = ORACLE_G * mass_at(masses, masses_are_f32, j) / r3;
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

/* force_mode: 0 = fast-math OpenMP loop (the reference's compiled kernel),
 *             1 = fast-math single thread (one generate_data.py worker),
 *             2 = strict ascending j, 3 = strict descending j. */
static void force(const double *pos, const void *masses, int masses_are_f32, int n,
                  double softening, int force_mode, double *acc)
{
    switch (force_mode) {
    case 0: oracle_accel_direct(pos, masses, masses_are_f32, n, softening, acc); break;
    case 1: oracle_accel_direct_serial(pos, masses, masses_are_f32, n, softening, acc); break;
    case 2: oracle_accel_direct_strict(pos, masses, masses_are_f32, n, softening, 0, acc); break;
    default: oracle_accel_direct_strict(pos, masses, masses_are_f32, n, softening, 1, acc); break;
    }
}

/* One NBodySimulator.step(), nbody.py:202-218, in place on pos/vel/acc (n,3). */
void oracle_step(double *pos, double *vel, double *acc, const void *masses,
                 int masses_are_f32, int n, double dt, double softening,
                 int force_mode)
{
    const double half_dt = 0.5 * dt;          /* "0.5 * self.dt" is evaluated first */
    const int m = 3 * n;
    for (int k = 0; k < m; ++k) vel[k] = vel[k] + half_dt * acc[k]; /* :205 */
    for (int k = 0; k < m; ++k) pos[k] = pos[k] + dt * vel[k];      /* :208 */
    force(pos, masses, masses_are_f32, n, softening, force_mode, acc); /* :211 */
    for (int k = 0; k < m; ++k) vel[k] = vel[k] + half_dt * acc[k]; /* :214 */
}

/*
 * NBodySimulator.run(n_steps, save_interval), nbody.py:220-248, without the
 * verbose energy print.  pos/vel/acc are the live state (updated in place).
 * Snapshots (get_state, :250-259) go to out_pos/out_vel/out_acc, each
 * (n_snap, n, 3) with n_snap = 1 + n_steps / save_interval, and out_time /
 * out_step (n_snap).  time0/step0 are the simulator's counters on entry;
 * the final counters are returned through time_out/step_out.  time is the
 * running float sum of dt (:217), not step*dt.
 */
void oracle_run(double *pos, double *vel, double *acc, const void *masses,
                int masses_are_f32, int n, double dt, double softening,
                int n_steps, int save_interval, int force_mode, double time0,
                long step0, double *out_pos, double *out_vel, double *out_acc,
                double *out_time, long *out_step, double *time_out, long *step_out)
{
    const size_t row = (size_t)3 * n;
    size_t s = 0;
    double t = time0;
    long step = step0;
    if (out_pos) memcpy(out_pos, pos, row * sizeof(double));  /* :235 */
    if (out_vel) memcpy(out_vel, vel, row * sizeof(double));
    if (out_acc) memcpy(out_acc, acc, row * sizeof(double));
    if (out_time) out_time[0] = t;
    if (out_step) out_step[0] = step;
    s = 1;
    for (int i = 0; i < n_steps; ++i) {                       /* :237 */
        oracle_step(pos, vel, acc, masses, masses_are_f32, n, dt, softening, force_mode);
        t = t + dt;                                           /* :217 */
        step += 1;                                            /* :218 */
        if ((i + 1) % save_interval == 0) {                   /* :240 */
            if (out_pos) memcpy(out_pos + s * row, pos, row * sizeof(double));
            if (out_vel) memcpy(out_vel + s * row, vel, row * sizeof(double));
            if (out_acc) memcpy(out_acc + s * row, acc, row * sizeof(double));
            if (out_time) out_time[s] = t;
            if (out_step) out_step[s] = step;
            ++s;
        }
    }
    if (time_out) *time_out = t;
    if (step_out) *step_out = step;
}

/*
 * The data-generation ensemble, generate_data.py:142-149: B independent
 * simulations, each advanced by one single-threaded worker (the reference pins
 * Numba to one thread per mp.Pool process, generate_data.py:16-19), workers in
 * parallel.  x0/v0 are (B,n,3); masses is shared (n) -- generate_data.py:108-109;
 * the initial accelerations are evaluated here from x0 and the shared masses
 * (generate_data.py:45-47).  Outputs are (B, n_snap, n, 3); any may be NULL.
 */
void oracle_ensemble_run(const double *x0, const double *v0, const void *masses,
                         int masses_are_f32, int B, int n, double dt,
                         double softening, int n_steps, int save_interval,
                         double *out_pos, double *out_vel, double *out_acc,
                         double *scratch /* B * 9n doubles */)
{
    const size_t row = (size_t)3 * n;
    const size_t n_snap = 1 + (size_t)(n_steps / save_interval);
#pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < B; ++b) {
        double *pos = scratch + (size_t)b * 3 * row;
        double *vel = pos + row;
        double *acc = vel + row;
        memcpy(pos, x0 + (size_t)b * row, row * sizeof(double));
        memcpy(vel, v0 + (size_t)b * row, row * sizeof(double));
        oracle_accel_direct_serial(pos, masses, masses_are_f32, n, softening, acc);
        oracle_run(pos, vel, acc, masses, masses_are_f32, n, dt, softening, n_steps,
                   save_interval, /*force_mode=*/1, 0.0, 0,
                   out_pos ? out_pos + (size_t)b * n_snap * row : NULL,
                   out_vel ? out_vel + (size_t)b * n_snap * row : NULL,
                   out_acc ? out_acc + (size_t)b * n_snap * row : NULL,
                   NULL, NULL, NULL, NULL);
    }
}
