// mtg_chain.cuh -- adjoints of the celerite coefficients and the chain rule back to the free log-parameters, shared by the
// two gradient kernels (K2: mtg_grad.cu, one thread per series; K11: mtg_scan_grad.cuh, one CTA per series).
#pragma once
#include "mtg_common.cuh"

// coefficient adjoints
template <int JR, int JC, bool DUAL>
struct CoefAdj {
    double ar[JR > 0 ? JR : 1], cr[JR > 0 ? JR : 1];
    double ac[JC > 0 ? JC : 1], bc[JC > 0 ? JC : 1], cc[JC > 0 ? JC : 1], dc[JC > 0 ? JC : 1];
    double a2[(DUAL && JC > 0) ? JC : 1], c2[(DUAL && JC > 0) ? JC : 1];
    double a_sum;
    __device__ __forceinline__ void zero()
    {
#pragma unroll
        for (int j = 0; j < JR; ++j) { ar[j] = 0.0; cr[j] = 0.0; }
#pragma unroll
        for (int j = 0; j < JC; ++j) { ac[j] = bc[j] = cc[j] = dc[j] = 0.0; if (DUAL) { a2[j] = 0.0; c2[j] = 0.0; } }
        a_sum = 0.0;
    }
};

// d loglike / d free parameter from the coefficient adjoints (chain rule through plan_coefficients)
template <int JR, int JC, bool DUAL>
__device__ __forceinline__ void plan_chain(const Plan &pl, const double *__restrict__ x,
                                           const Coef<JR, JC, DUAL> &c, const CoefAdj<JR, JC, DUAL> &g,
                                           double *__restrict__ grad /* [n_free], pre-zeroed */)
{
    auto add = [&](int full_idx, double v) { const int pos = pl.src_pos[full_idx]; if (pos >= 0) grad[pos] += v; };
    for (int k = 0; k < pl.n_jit; ++k)
        add(pl.jit_idx[k], pl.jit_raw[k] ? g.a_sum : g.a_sum * 2.0 * mtg_exp(2.0 * plan_param(pl, x, pl.jit_idx[k])));
#pragma unroll
    for (int j = 0; j < JR; ++j) {
        add(pl.r_ia[j], (g.ar[j] + g.a_sum) * (pl.r_raw[j] ? 1.0 : c.ar[j]));
        add(pl.r_ic[j], g.cr[j] * (pl.r_raw[j] ? 1.0 : -c.ncr[j] * MTG_LN2));
    }
#pragma unroll
    for (int j = 0; j < JC; ++j) {
        const double a = c.ac[j], b = c.bc[j], cc = -c.ncc[j] * MTG_LN2, d = c.dc[j];
        const double ga = g.ac[j] + g.a_sum, gb = g.bc[j], gc = g.cc[j], gd = g.dc[j];
        const int *ix = pl.c_idx[j];
        switch (pl.c_kind[j]) {
        case CS_COMPLEX3: add(ix[0], ga * a); add(ix[1], gc * cc); add(ix[2], gd * d); break;
        case CS_COMPLEX4: add(ix[0], ga * a); add(ix[1], gb * b); add(ix[2], gc * cc); add(ix[3], gd * d); break;
        case CS_RAW: add(ix[0], ga); add(ix[1], gb); add(ix[2], gc); add(ix[3], gd); break;
        case CS_LORENTZIAN: add(ix[0], ga * a); add(ix[1], -gc * cc); add(ix[2], gc * cc + gd * d); break;
        case CS_COSINUS: add(ix[0], ga * a); add(ix[1], gd * d); break;
        case CS_BENDINGPL: add(ix[0], ga * a); add(ix[1], gb * b); add(ix[2], (gc + gd) * d); break;
        case CS_MATERN32: add(ix[0], 2.0 * (ga * a + gb * b)); add(ix[1], -(gb * b + gc * cc)); break;
        default: /* CS_SHO */ {
            const double Q = mtg_exp(plan_param(pl, x, ix[1]));
            if (!DUAL || c.pair[j]) {
                const double f2 = 4.0 * Q * Q - 1.0, r = 4.0 * Q * Q / f2;
                add(ix[0], ga * a + gb * b);
                add(ix[1], ga * a + gb * b * (1.0 - r) - gc * cc + gd * d * (r - 1.0));
                add(ix[2], ga * a + gb * b + gc * cc + gd * d);
            } else {
                const double a2 = c.a2[j], c2 = -c.nc2[j] * MTG_LN2;
                const double ga2 = g.a2[j] + g.a_sum, gc2 = g.c2[j];
                const double f = sqrt(1.0 - 4.0 * Q * Q);
                const double h = 0.5 * (a + a2);                       // = 1/2 S0 w0 Q
                const double da = h * 4.0 * Q * Q / (f * f * f);
                const double w0 = mtg_exp(plan_param(pl, x, ix[2]));
                const double dcq = 2.0 * w0 * Q / f;
                add(ix[0], ga * a + ga2 * a2);
                add(ix[1], ga * (a + da) + ga2 * (a2 - da) + gc * (-cc + dcq) + gc2 * (-c2 - dcq));
                add(ix[2], ga * a + ga2 * a2 + gc * cc + gc2 * c2);
            } } break;
        }
    }
}

