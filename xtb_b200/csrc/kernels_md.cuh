// Molecular dynamics propagation: one leap-frog step with Berendsen thermostat, SHAKE bond constraints
// and removal of the overall translation / rotation, for every molecule of the batch (independent replicas).  Positions,
// velocities and forces stay on the device between single points; the only traffic per step is the energy read-back.
//
// Reference being replaced: src/dynamic.f90:140-702 (md: steps 1-3 at 553-626), src/shake_module.f90:150-245 (do_shake),
// src/rmrottr.f90:19-125 (rmrottr, centerofmass), src/dynamic.f90:767-776 (ekinet).
#pragma once
#include "common.cuh"

namespace xtbb200 {

struct MdDev {
    double *velo, *vel, *mass, *xyzo, *acc, *xyzt;   // per atom (3 per atom except mass)
    // SHAKE: constraints of all molecules concatenated, cons0[nmol + 1] ranges; atoms are molecule-local indices
    const long long* cons0;
    const int* cons;        // [2] per constraint
    const double* dist2;    // squared reference distance
    double *dro, *gc;       // [3] / [1] per constraint
    const long long* adj0;  // per atom of the batch (+1): range in adj
    const int* adj;         // constraint index within the molecule * 2 + side (0: first atom, 1: second), ascending
    const double* nfree;    // degrees of freedom per molecule
    double* out;            // [4] per molecule: epot, ekin, T, SHAKE cycles (negative: not converged)
    double tstep, tstep0, taut, tinit;
    int thermostat, shake;
};

constexpr int kShakeMaxCyc = 250;        // src/shake_module.f90:33-34
constexpr double kShakeTol = 1.0e-7;

__device__ __forceinline__ double block_max(double v, double* red) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_max(v);
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    double r = red[0];
    for (int i = 1; i < nw; ++i) r = fmax(r, red[i]);
    return r;
}

__global__ void __launch_bounds__(256) k_md_update(Dev d, MdDev md) {
    const int im = blockIdx.x;
    const MolDesc m = d.mol[im];
    __shared__ double red[32];
    const int tid = threadIdx.x, nt = blockDim.x;
    const double tstep = md.tstep;
    const double nfree = md.nfree[im];
    // ---- acceleration at t, estimate of the velocities at t, temperature (dynamic.f90:553-575)
    double ek = 0;
    for (int i = tid; i < m.nat; i += nt) {
        const int a = m.at0 + i;
        const double ms = md.mass[a];
        double v2 = 0;
        for (int k = 0; k < 3; ++k) {
            const double ac = -d.grad[3 * a + k] / ms;
            md.acc[3 * a + k] = ac;
            md.xyzo[3 * a + k] = d.xyz[3 * a + k];
            const double vn = md.velo[3 * a + k] + 0.5 * tstep * ac;
            v2 += vn * vn;
        }
        ek += ms * v2;
    }
    double ekin = 0.5 * block_sum(ek, red);
    double T = 2.0 * ekin / nfree / kBoltz;
    const double xlam2 = T > 0.0 ? sqrt(1.0 + (md.tstep0 / md.taut) * (md.tinit / T - 1.0)) : 1.0;   // start from rest: no scaling
    // ---- velocity and position update (dynamic.f90:577-600)
    ek = 0;
    for (int i = tid; i < m.nat; i += nt) {
        const int a = m.at0 + i;
        double v2 = 0;
        for (int k = 0; k < 3; ++k) {
            const double vo = md.velo[3 * a + k];
            double v = vo + md.acc[3 * a + k] * tstep;
            if (md.thermostat) v = xlam2 * v;
            md.vel[3 * a + k] = v;
            const double x = md.xyzo[3 * a + k] + v * tstep;
            d.xyz[3 * a + k] = x;
            md.xyzt[3 * a + k] = x;
            const double vn = 0.5 * (vo + v);
            v2 += vn * vn;
        }
        ek += md.mass[a] * v2;
    }
    ekin = 0.5 * block_sum(ek, red);
    T = 2.0 * ekin / nfree / kBoltz;
    // ---- SHAKE at t + dt (do_shake): every cycle evaluates all constraints on the same positions, then corrects them
    double cycles = 0;
    if (md.shake) {
        const long long c0 = md.cons0[im];
        const int nc = (int)(md.cons0[im + 1] - c0);
        if (nc > 0) {
            for (int c = tid; c < nc; c += nt) {
                const int ia = m.at0 + md.cons[2 * (c0 + c)], ja = m.at0 + md.cons[2 * (c0 + c) + 1];
                for (int k = 0; k < 3; ++k) md.dro[3 * (c0 + c) + k] = md.xyzo[3 * ia + k] - md.xyzo[3 * ja + k];
            }
            __syncthreads();
            bool conv = false;
            int icyc = 0;
            do {
                double mx = 0;
                for (int c = tid; c < nc; c += nt) {
                    const int ia = m.at0 + md.cons[2 * (c0 + c)], ja = m.at0 + md.cons[2 * (c0 + c) + 1];
                    const double dx = md.xyzt[3 * ia] - md.xyzt[3 * ja], dy = md.xyzt[3 * ia + 1] - md.xyzt[3 * ja + 1],
                                 dz = md.xyzt[3 * ia + 2] - md.xyzt[3 * ja + 2];
                    const double d2 = dx * dx + dy * dy + dz * dz, dist = md.dist2[c0 + c];
                    mx = fmax(mx, fabs(d2 - dist) / dist);
                    const double rmi = 1.0 / md.mass[ia], rmj = 1.0 / md.mass[ja];
                    const double* o = md.dro + 3 * (c0 + c);
                    const double denom = 2.0 * (rmi + rmj) * (dx * o[0] + dy * o[1] + dz * o[2]);
                    md.gc[c0 + c] = (dist - d2) / denom;
                }
                mx = block_max(mx, red);   // barriers: gc is visible to the whole CTA
                conv = mx < kShakeTol;
                if (!conv) {
                    for (int i = tid; i < m.nat; i += nt) {
                        const int a = m.at0 + i;
                        const double rm = 1.0 / md.mass[a];
                        double x = md.xyzt[3 * a], y = md.xyzt[3 * a + 1], z = md.xyzt[3 * a + 2];
                        for (long long e = md.adj0[a]; e < md.adj0[a + 1]; ++e) {
                            const int c = md.adj[e] >> 1;
                            const double s = (md.adj[e] & 1) ? -1.0 : 1.0;
                            const double g = s * rm * md.gc[c0 + c];
                            const double* o = md.dro + 3 * (c0 + c);
                            x += g * o[0]; y += g * o[1]; z += g * o[2];
                        }
                        md.xyzt[3 * a] = x; md.xyzt[3 * a + 1] = y; md.xyzt[3 * a + 2] = z;
                    }
                }
                __syncthreads();
                icyc++;
            } while (!conv && icyc <= kShakeMaxCyc);
            cycles = conv ? (double)icyc : -(double)icyc;
            if (conv) {
                const double tau1 = 1.0 / tstep, tau2 = tau1 * tau1;
                for (int i = tid; i < m.nat; i += nt) {
                    const int a = m.at0 + i;
                    for (int k = 0; k < 3; ++k) {
                        const double dx = md.xyzt[3 * a + k] - d.xyz[3 * a + k];
                        md.vel[3 * a + k] += dx * tau1;
                        md.acc[3 * a + k] += dx * tau2;
                        d.xyz[3 * a + k] = md.xyzt[3 * a + k];
                    }
                }
            }
            __syncthreads();
        }
    }
    // ---- velo = vel, remove translation and rotation (rmrottr)
    double p[4] = {0, 0, 0, 0};
    for (int i = tid; i < m.nat; i += nt) {
        const int a = m.at0 + i;
        const double ms = md.mass[a];
        p[0] += ms;
        for (int k = 0; k < 3; ++k) p[1 + k] += ms * d.xyz[3 * a + k];
    }
    const double tmass = block_sum(p[0], red);
    const double cx = block_sum(p[1], red) / tmass, cy = block_sum(p[2], red) / tmass, cz = block_sum(p[3], red) / tmass;
    double q[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};   // angmom(3), ixx iyy izz ixy ixz iyz, rlm(3)
    for (int i = tid; i < m.nat; i += nt) {
        const int a = m.at0 + i;
        const double ms = md.mass[a];
        const double x = d.xyz[3 * a] - cx, y = d.xyz[3 * a + 1] - cy, z = d.xyz[3 * a + 2] - cz;
        const double vx = md.vel[3 * a], vy = md.vel[3 * a + 1], vz = md.vel[3 * a + 2];
        q[0] += ms * (y * vz - z * vy); q[1] += ms * (z * vx - x * vz); q[2] += ms * (x * vy - y * vx);
        q[3] += ms * (y * y + z * z); q[4] += ms * (z * z + x * x); q[5] += ms * (x * x + y * y);
        q[6] -= ms * x * y; q[7] -= ms * x * z; q[8] -= ms * y * z;
        q[9] += ms * vx; q[10] += ms * vy; q[11] += ms * vz;
    }
    for (int k = 0; k < 12; ++k) q[k] = block_sum(q[k], red);
    // omega = I^-1 L (symmetric 3 x 3 inverse by cofactors)
    const double ixx = q[3], iyy = q[4], izz = q[5], ixy = q[6], ixz = q[7], iyz = q[8];
    const double c00 = iyy * izz - iyz * iyz, c01 = ixz * iyz - ixy * izz, c02 = ixy * iyz - ixz * iyy;
    const double c11 = ixx * izz - ixz * ixz, c12 = ixy * ixz - ixx * iyz, c22 = ixx * iyy - ixy * ixy;
    const double det = ixx * c00 + ixy * c01 + ixz * c02;
    // atoms, diatomics and exactly linear molecules have a singular inertia tensor: only the translation is removed
    const double scale = ixx + iyy + izz, rdet = fabs(det) > 1.0e-12 * scale * scale * scale ? 1.0 / det : 0.0;
    const double ox = (c00 * q[0] + c01 * q[1] + c02 * q[2]) * rdet, oy = (c01 * q[0] + c11 * q[1] + c12 * q[2]) * rdet,
                 oz = (c02 * q[0] + c12 * q[1] + c22 * q[2]) * rdet;
    for (int i = tid; i < m.nat; i += nt) {
        const int a = m.at0 + i;
        const double x = d.xyz[3 * a] - cx, y = d.xyz[3 * a + 1] - cy, z = d.xyz[3 * a + 2] - cz;
        md.velo[3 * a] = md.vel[3 * a] - q[9] / tmass - (oy * z - oz * y);
        md.velo[3 * a + 1] = md.vel[3 * a + 1] - q[10] / tmass - (oz * x - ox * z);
        md.velo[3 * a + 2] = md.vel[3 * a + 2] - q[11] / tmass - (ox * y - oy * x);
    }
    if (tid == 0) {
        double* o = md.out + 4 * im;
        o[0] = d.st[im].energy; o[1] = ekin; o[2] = T; o[3] = cycles;
    }
}

}  // namespace xtbb200
