// Stage 1 kernels: batched HEALPix ring transforms.  See sht.cuh for the reference functions replaced.
//
// Every step is a real GEMM whose N dimension is the shell batch, run on DMMA:
//   ring analysis   F_m(ring)[c]   = Σ_j  e^{-imφ_j} f_j[shell]          A = twiddles (generated), B = map
//   Legendre anal.  a_lm[c]        = w Σ_k λ_lm(θ_k) (F_N ± F_S)_m[k][c]  A = λ table,  B = F (parity-combined)
//   Legendre synth. G_m(ring)[c]   = Σ_l λ_lm(θ_k) a_lm[c]  (E ± O)       A = λᵀ,       B = alm
//   ring synthesis  f_j[shell]     = Σ_m (2-δ_m0) Re(G_m e^{imφ_j})       A = twiddles, B = G   (+ residual map - f)
// c = comp*nrp + shell (re plane, then im plane).
#include "sht.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>

namespace sfb {

constexpr int kT = 256;      // threads per CTA (8 warps as 4 x 2 over a 64 x 64 tile)
constexpr int kLdA = 36;     // 32 + 4  (≡ 4 mod 16)
constexpr int kLdB = 68;     // 64 + 4  (≡ 4 mod 16)

__device__ __forceinline__ size_t lm_mmajor(int lmax, int l, int m) {
    return (size_t)l + ((size_t)m * (2 * lmax + 1 - m)) / 2;
}

// ---------------------------------------------------------------------------------------------
// setup kernels

__global__ void twiddle_table_kernel(double2* __restrict__ tw, int nside) {
    // block i-1 of the table: ring length nφ = 4i, entries t in [0, 8i) at offset 4i(i-1)
    const int i = blockIdx.x + 1;
    double2* dst = tw + (size_t)4 * i * (i - 1);
    for (int t = threadIdx.x; t < 8 * i; t += blockDim.x) {
        double s, c;
        sincospi((double)t / (double)(4 * i), &s, &c);
        dst[t] = make_double2(c, s);
    }
}

__device__ __forceinline__ void ring_z_sth(int nside, int k /*0-based north ring*/, double& z, double& sth) {
    const int i = k + 1;
    if (i < nside) {
        const double omz = (double)i * i / (3.0 * nside * nside);
        z = 1.0 - omz;
        sth = sqrt(omz * (1.0 + z));
    } else {
        z = (2 * nside - i) * 2.0 / (3.0 * nside);
        sth = sqrt((1.0 - z) * (1.0 + z));
    }
}

// λ_lm(θ_k) for the northern rings (incl. equator); thread = (ring k, m), sequential three-term recurrence in l.
__global__ void lambda_table_kernel(double* __restrict__ lam, int nside, int lmax, int nhalf) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int m = blockIdx.y;
    if (k >= nhalf) return;
    double z, sth;
    ring_z_sth(nside, k, z, sth);
    double lmm = 0.28209479177387814;  // sqrt(1/(4π))
    for (int mm = 1; mm <= m; ++mm) lmm = -lmm * sth * sqrt((2.0 * mm + 1.0) / (2.0 * mm));
    lam[lm_mmajor(lmax, m, m) * nhalf + k] = lmm;
    if (m == lmax) return;
    double l2 = lmm, l1 = z * sqrt(2.0 * m + 3.0) * lmm;
    lam[lm_mmajor(lmax, m + 1, m) * nhalf + k] = l1;
    const double m2 = (double)m * m;
    for (int l = m + 2; l <= lmax; ++l) {
        const double a = sqrt((4.0 * l * l - 1.0) / ((double)l * l - m2));
        const double b = sqrt(((l - 1.0) * (l - 1.0) - m2) / (4.0 * (l - 1.0) * (l - 1.0) - 1.0));
        const double ln = a * (z * l1 - b * l2);
        lam[lm_mmajor(lmax, l, m) * nhalf + k] = ln;
        l2 = l1;
        l1 = ln;
    }
}

// max_l |λ_lm(θ_k)| per (north ring k, m): input of the m-cutoff tables (libsharp's `mlim` idea, but read off the table the
// kernels actually use, so the cutoff is rigorous): beyond mlim[k] every λ_lm(θ_k), l <= lmax, is below 1e-30 and the
// (ring, m) pair contributes nothing at FP64 precision.
__global__ void lambda_absmax_kernel(const double* __restrict__ lam, int lmax, int nhalf, double* __restrict__ out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int m = blockIdx.y;
    if (k >= nhalf) return;
    double mx = 0.0;
    for (int l = m; l <= lmax; ++l) mx = fmax(mx, fabs(lam[lm_mmajor(lmax, l, m) * nhalf + k]));
    out[(size_t)m * nhalf + k] = mx;
}

// ---------------------------------------------------------------------------------------------
// udgrade via the NESTED scheme

__device__ __constant__ int c_jrll[12] = {2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4};
__device__ __constant__ int c_jpll[12] = {1, 3, 5, 7, 0, 2, 4, 6, 1, 3, 5, 7};

__device__ __forceinline__ unsigned spread_bits(unsigned v) {
    v = (v | (v << 8)) & 0x00FF00FFu;
    v = (v | (v << 4)) & 0x0F0F0F0Fu;
    v = (v | (v << 2)) & 0x33333333u;
    v = (v | (v << 1)) & 0x55555555u;
    return v;
}
__device__ __forceinline__ unsigned compress_bits(unsigned v) {
    v &= 0x55555555u;
    v = (v | (v >> 1)) & 0x33333333u;
    v = (v | (v >> 2)) & 0x0F0F0F0Fu;
    v = (v | (v >> 4)) & 0x00FF00FFu;
    v = (v | (v >> 8)) & 0x0000FFFFu;
    return v;
}
__device__ __forceinline__ int isqrt_ll(long long v) {
    long long r = (long long)sqrt((double)v + 0.5);
    while (r * r > v) --r;
    while ((r + 1) * (r + 1) <= v) ++r;
    return (int)r;
}

__device__ long long ring2nest_dev(int nside, long long pix) {
    const long long ncap = 2LL * nside * (nside - 1), npix = 12LL * nside * nside;
    const int nl2 = 2 * nside;
    int iring, iphi, kshift, nr, face;
    if (pix < ncap) {
        iring = (1 + isqrt_ll(1 + 2 * pix)) >> 1;
        iphi = (int)(pix + 1 - 2LL * iring * (iring - 1));
        kshift = 0;
        nr = iring;
        face = (iphi - 1) / nr;
    } else if (pix < npix - ncap) {
        const long long ip = pix - ncap;
        const int tmp = (int)(ip / (4 * nside));
        iring = tmp + nside;
        iphi = (int)(ip - (long long)tmp * 4 * nside) + 1;
        kshift = (iring + nside) & 1;
        nr = nside;
        const int ire = tmp + 1, irm = nl2 + 1 - tmp;
        const int ifm = (iphi - ire / 2 + nside - 1) / nside;
        const int ifp = (iphi - irm / 2 + nside - 1) / nside;
        face = (ifp == ifm) ? (ifp | 4) : ((ifp < ifm) ? ifp : (ifm + 8));
    } else {
        const long long ip = npix - pix;
        iring = (1 + isqrt_ll(2 * ip - 1)) >> 1;
        iphi = 4 * iring + 1 - (int)(ip - 2LL * iring * (iring - 1));
        kshift = 0;
        nr = iring;
        iring = 2 * nl2 - iring;
        face = 8 + (iphi - 1) / nr;
    }
    const int irt = iring - c_jrll[face] * nside + 1;
    int ipt = 2 * iphi - c_jpll[face] * nr - kshift - 1;
    if (ipt >= nl2) ipt -= 8 * nside;
    const int ix = (ipt - irt) >> 1, iy = (-ipt - irt) >> 1;
    return (long long)face * nside * nside + (spread_bits((unsigned)ix) | (spread_bits((unsigned)iy) << 1));
}

__device__ long long nest2ring_dev(int nside, long long ipnest) {
    const long long npface = (long long)nside * nside, npix = 12 * npface, ncap = 2LL * nside * (nside - 1);
    const int nl4 = 4 * nside;
    const int face = (int)(ipnest / npface);
    const unsigned ipf = (unsigned)(ipnest % npface);
    const int ix = (int)compress_bits(ipf), iy = (int)compress_bits(ipf >> 1);
    const int jrt = ix + iy, jpt = ix - iy;
    const int jr = c_jrll[face] * nside - jrt - 1;
    int nr, kshift;
    long long n_before;
    if (jr < nside) {
        nr = jr;
        n_before = 2LL * nr * (nr - 1);
        kshift = 0;
    } else if (jr > 3 * nside) {
        nr = nl4 - jr;
        n_before = npix - 2LL * (nr + 1) * nr;
        kshift = 0;
    } else {
        nr = nside;
        n_before = ncap + (long long)(jr - nside) * nl4;
        kshift = (jr - nside) & 1;
    }
    int jp = (c_jpll[face] * nr + jpt + 1 + kshift) / 2;
    if (jp > nl4) jp -= nl4;
    if (jp < 1) jp += nl4;
    return n_before + jp - 1;
}

// out[pix_out][shell] (stride nrp, padded shells zeroed) from in[pix_in][shell] (stride ldw)
__global__ void udgrade_kernel(const double* __restrict__ in, long long ldw, int nside_in, double* __restrict__ out,
                               int nside_out, int nr, int nrp) {
    const long long npix_out = 12LL * nside_out * nside_out;
    const long long total = npix_out * nrp;
    for (long long x = blockIdx.x * (long long)blockDim.x + threadIdx.x; x < total;
         x += (long long)gridDim.x * blockDim.x) {
        const long long po = x / nrp;
        const int sh = (int)(x - po * nrp);
        double v = 0.0;
        if (sh < nr) {
            const long long nest_o = ring2nest_dev(nside_out, po);
            if (nside_out >= nside_in) {
                const long long ratio = (long long)(nside_out / nside_in) * (nside_out / nside_in);
                v = in[nest2ring_dev(nside_in, nest_o / ratio) * ldw + sh];
            } else {
                const long long ratio = (long long)(nside_in / nside_out) * (nside_in / nside_out);
                double s = 0.0;
                for (long long c = 0; c < ratio; ++c) s += in[nest2ring_dev(nside_in, nest_o * ratio + c) * ldw + sh];
                v = s / (double)ratio;
            }
        }
        out[x] = v;
    }
}

// ---------------------------------------------------------------------------------------------
struct RingTabs {
    const int* nphi;
    const int* start;
    const int* shift;
    const int* twoff;
    const double2* tw;
};

// F[m][ring][c] = Σ_j e^{-imφ_j} f_j      CTA = (ring, 32 m's, 64 shells); rows 0-31 -> Re, 32-63 -> Im
__global__ void __launch_bounds__(kT) ring_analysis_kernel(const double* __restrict__ map, long long ldw, int nr, int nrp,
                                                           RingTabs rt, const int* __restrict__ ring_list, int nrings,
                                                           int lmax, double* __restrict__ F) {
    __shared__ double As[64 * kLdA];
    __shared__ double Bs[32 * kLdB];
    const int ring = ring_list[blockIdx.x], m0 = blockIdx.y * 32, sh0 = blockIdx.z * 64;
    const int nphi = rt.nphi[ring], start = rt.start[ring], s = rt.shift[ring];
    const double2* tw = rt.tw + rt.twoff[ring];
    const unsigned two_nphi = 2u * nphi;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int wm = warp >> 1, wn = warp & 1;
    double acc[2][4][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    for (int j0 = 0; j0 < nphi; j0 += 32) {
        {
            const int kk = tid & 31, j = j0 + kk;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int mrow = (tid >> 5) + 8 * q, m = m0 + mrow;
                double c = 0.0, sn = 0.0;
                if (j < nphi && m <= lmax) {
                    const unsigned tt = ((unsigned)m * (unsigned)(2 * j + s)) % two_nphi;
                    const double2 w = tw[tt];
                    c = w.x;
                    sn = -w.y;
                }
                As[mrow * kLdA + kk] = c;
                As[(mrow + 32) * kLdA + kk] = sn;
            }
            const int c = tid & 63;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int k2 = (tid >> 6) + 4 * q;
                double v = 0.0;
                if (j0 + k2 < nphi && sh0 + c < nr) v = map[(size_t)(start + j0 + k2) * ldw + sh0 + c];
                Bs[k2 * kLdB + c] = v;
            }
        }
        __syncthreads();
        warp_gemm_ss<2, 4>(acc, As + wm * 16 * kLdA, kLdA, Bs + wn * 32, kLdB, 32);
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int row = wm * 16 + i * 8 + g;
        const int m = m0 + (row & 31), comp = row >> 5;
        if (m > lmax) continue;
        double* dst = F + ((size_t)m * nrings + ring) * 2 * nrp + (size_t)comp * nrp + sh0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int col = wn * 32 + j * 8 + 2 * t;
            if (sh0 + col < nrp) {
                dst[col] = acc[i][j][0];
                dst[col + 1] = acc[i][j][1];
            }
        }
    }
}

// f_j = Σ_m (2-δ_m0) (Re G_m cos mφ_j - Im G_m sin mφ_j);  out = residual ? map - f : f
// CTA = (tile = (ring, 64 pixels), 64 shells)
__global__ void __launch_bounds__(kT) ring_synthesis_kernel(const double* __restrict__ G, RingTabs rt,
                                                            const int* __restrict__ tile_ring,
                                                            const int* __restrict__ tile_j0, int nrings, int lmax,
                                                            int nr, int nrp, const double* __restrict__ map,
                                                            long long ldw, int residual, double* __restrict__ out) {
    __shared__ double As[64 * kLdA];
    __shared__ double Bs[32 * kLdB];
    const int ring = tile_ring[blockIdx.x], j0 = tile_j0[blockIdx.x], sh0 = blockIdx.y * 64;
    const int nphi = rt.nphi[ring], start = rt.start[ring], s = rt.shift[ring];
    const double2* tw = rt.tw + rt.twoff[ring];
    const unsigned two_nphi = 2u * nphi;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int wm = warp >> 1, wn = warp & 1;
    double acc[2][4][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    for (int m0 = 0; m0 <= lmax; m0 += 16) {
        {
            const int row = tid & 63, j = j0 + row;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int mm = (tid >> 6) + 4 * q, m = m0 + mm;
                double c = 0.0, sn = 0.0;
                if (j < nphi && m <= lmax) {
                    const unsigned tt = ((unsigned)m * (unsigned)(2 * j + s)) % two_nphi;
                    const double2 w = tw[tt];
                    const double cm = (m == 0) ? 1.0 : 2.0;
                    c = cm * w.x;
                    sn = -cm * w.y;
                }
                As[row * kLdA + mm] = c;
                As[row * kLdA + 16 + mm] = sn;
            }
            const int c = tid & 63;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int k2 = (tid >> 6) + 4 * q;
                const int m = m0 + (k2 & 15), part = k2 >> 4;
                double v = 0.0;
                if (m <= lmax && sh0 + c < nrp) v = G[((size_t)m * nrings + ring) * 2 * nrp + (size_t)part * nrp + sh0 + c];
                Bs[k2 * kLdB + c] = v;
            }
        }
        __syncthreads();
        warp_gemm_ss<2, 4>(acc, As + wm * 16 * kLdA, kLdA, Bs + wn * 32, kLdB, 32);
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int j = j0 + wm * 16 + i * 8 + g;
        if (j >= nphi) continue;
        const size_t pix = (size_t)start + j;
#pragma unroll
        for (int jn = 0; jn < 4; ++jn) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int sh = sh0 + wn * 32 + jn * 8 + 2 * t + e;
                if (sh >= nrp) continue;
                double v = 0.0;
                if (sh < nr) {
                    v = acc[i][jn][e];
                    if (residual) v = map[pix * ldw + sh] - v;
                }
                out[pix * nrp + sh] = v;
            }
        }
    }
}


// ---------------------------------------------------------------------------------------------
// Equatorial-belt rings (nφ = 4 nside, a power of two): shared-memory radix-2 FFT, two real shells packed into
// one complex sequence.  Z holds n x SCH complex values; input in bit-reversed order, output in natural order.
// R-point DFT of a register array (decimation in frequency, then the bit-reversal is undone by register renaming).
__device__ __constant__ double c_cos16[8] = {1.0, 0.92387953251128674, 0.70710678118654752, 0.38268343236508977,
                                            0.0, -0.38268343236508977, -0.70710678118654752, -0.92387953251128674};
__device__ __constant__ double c_sin16[8] = {0.0, 0.38268343236508977, 0.70710678118654752, 0.92387953251128674,
                                            1.0, 0.92387953251128674, 0.70710678118654752, 0.38268343236508977};

template <int R, int DIR>
__device__ __forceinline__ void dft_reg(double2 (&x)[R]) {
#pragma unroll
    for (int h = R / 2; h >= 1; h >>= 1) {
#pragma unroll
        for (int blk = 0; blk < R; blk += 2 * h) {
#pragma unroll
            for (int k = 0; k < h; ++k) {
                const double2 u = x[blk + k], v = x[blk + k + h];
                x[blk + k] = make_double2(u.x + v.x, u.y + v.y);
                const double dx = u.x - v.x, dy = u.y - v.y;
                const int ti = k * (8 / h);  // W_{2h}^k = W_16^{k * 8/h}
                if (ti == 0) {
                    x[blk + k + h] = make_double2(dx, dy);
                } else if (ti == 4) {       // multiply by (0, DIR) i.e. ±i ... W = e^{DIR iπ/2}
                    x[blk + k + h] = make_double2(-DIR * dy, DIR * dx);
                } else {
                    const double wr = c_cos16[ti], wi = DIR * c_sin16[ti];
                    x[blk + k + h] = make_double2(dx * wr - dy * wi, dx * wi + dy * wr);
                }
            }
        }
    }
    // undo the bit reversal (compile-time permutation)
    double2 y[R];
#pragma unroll
    for (int i = 0; i < R; ++i) {
        int r = 0;
#pragma unroll
        for (int b = 1, c = R >> 1; b < R; b <<= 1, c >>= 1)
            if (i & b) r |= c;
        y[r] = x[i];
    }
#pragma unroll
    for (int i = 0; i < R; ++i) x[i] = y[i];
}

// One Stockham (autosort) radix-R pass over Z[n][sch] in shared memory, in place: every thread reads the inputs of
// its butterflies into registers, barrier, writes the outputs, barrier.  Ns = product of the previous radices.
// tw[idx] = (cos, sin)(π idx / n), idx in [0, 2n).  At most 16 complex values per thread (n * sch <= 4096).
template <int R, int DIR>
__device__ __forceinline__ void stockham_pass(double2* Z, int n, int sch, int Ns, const double2* __restrict__ tw) {
    constexpr int MAXU = 16 / R;
    const int nb = n / R, nbf = nb * sch;
    const int tstep = (2 * n) / (Ns * R);
    double2 x[MAXU][R];
    int j0s[MAXU];
#pragma unroll
    for (int u = 0; u < MAXU; ++u) {
        const int b = threadIdx.x + u * kT;
        if (b < nbf) {
            const int lane = b % sch, j = b / sch, k = j & (Ns - 1);
#pragma unroll
            for (int t = 0; t < R; ++t) x[u][t] = Z[(j + t * nb) * sch + lane];
#pragma unroll
            for (int t = 1; t < R; ++t) {
                const double2 w = tw[k * t * tstep];
                const double wr = w.x, wi = DIR * w.y;
                const double2 v = x[u][t];
                x[u][t] = make_double2(v.x * wr - v.y * wi, v.x * wi + v.y * wr);
            }
            dft_reg<R, DIR>(x[u]);
            j0s[u] = ((j - k) * R + k) * sch + lane;
        }
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < MAXU; ++u) {
        const int b = threadIdx.x + u * kT;
        if (b < nbf) {
#pragma unroll
            for (int t = 0; t < R; ++t) Z[j0s[u] + t * Ns * sch] = x[u][t];
        }
    }
    __syncthreads();
}

// n-point FFT (n = 2^log2n) of the sch interleaved sequences in Z, natural order in and out.
template <int DIR>  // -1: forward e^{-iθ}, +1: backward e^{+iθ}
__device__ __forceinline__ void fft_smem(double2* Z, int n, int log2n, int sch, const double2* __restrict__ tw) {
    int Ns = 1, rem = log2n;
    while (rem >= 4) {
        stockham_pass<16, DIR>(Z, n, sch, Ns, tw);
        Ns *= 16;
        rem -= 4;
    }
    if (rem == 3)
        stockham_pass<8, DIR>(Z, n, sch, Ns, tw);
    else if (rem == 2)
        stockham_pass<4, DIR>(Z, n, sch, Ns, tw);
    else if (rem == 1)
        stockham_pass<2, DIR>(Z, n, sch, Ns, tw);
}

// F_m = e^{-imφ0} X[m mod n] for one belt ring and 2*sch shells     CTA = (belt ring, shell chunk)
__global__ void __launch_bounds__(kT) belt_analysis_fft_kernel(const double* __restrict__ map, long long ldw, int nr,
                                                               int nrp, RingTabs rt, const int* __restrict__ ring_list,
                                                               int nrings, int lmax, int log2n, int sch,
                                                               double* __restrict__ F) {
    extern __shared__ double2 Zs[];
    const int ring = ring_list[blockIdx.x], sh0 = blockIdx.y * 2 * sch;
    const int n = rt.nphi[ring], start = rt.start[ring], s = rt.shift[ring];
    const double2* tw = rt.tw + rt.twoff[ring];
    const int nsh = 2 * sch;
    for (int x = threadIdx.x; x < n * sch; x += blockDim.x) {
        const int lane = x % sch, j = x / sch;
        const int sa = sh0 + 2 * lane;
        const double* src = map + (size_t)(start + j) * ldw + sa;
        const double a = (sa < nr) ? src[0] : 0.0, b = (sa + 1 < nr) ? src[1] : 0.0;
        Zs[j * sch + lane] = make_double2(a, b);
    }
    __syncthreads();
    fft_smem<-1>(Zs, n, log2n, sch, tw);
    for (int x = threadIdx.x; x <= lmax * sch + sch - 1; x += blockDim.x) {
        const int lane = x % sch, m = x / sch;
        const int k = m % n, kc = (n - k) % n;
        const double2 z1 = Zs[k * sch + lane], z2 = Zs[kc * sch + lane];
        // X_a = (z1 + conj z2)/2, X_b = (z1 - conj z2)/(2i)
        const double ar = 0.5 * (z1.x + z2.x), ai = 0.5 * (z1.y - z2.y);
        const double br = 0.5 * (z1.y + z2.y), bi = -0.5 * (z1.x - z2.x);
        const double2 w = tw[(unsigned)(m * s) % (2u * n)];  // e^{-imφ0} = (cos, -sin)(π m s / n)
        const double c = w.x, sn = -w.y;
        double* dst = F + ((size_t)m * nrings + ring) * 2 * nrp + sh0 + 2 * lane;
        if (sh0 + 2 * lane < nrp) {
            dst[0] = ar * c - ai * sn;
            dst[1] = br * c - bi * sn;
            dst[nrp] = ar * sn + ai * c;
            dst[nrp + 1] = br * sn + bi * c;
        }
    }
    (void)nsh;
}

// f_j = Σ_m (2-δ_m0) Re(G_m e^{imφ_j}) for one belt ring and 2*sch shells, out = residual ? map - f : f
__global__ void __launch_bounds__(kT) belt_synthesis_fft_kernel(const double* __restrict__ G, RingTabs rt,
                                                                const int* __restrict__ ring_list, int nrings, int lmax,
                                                                int log2n, int sch, int nr, int nrp,
                                                                const double* __restrict__ map, long long ldw,
                                                                int residual, double* __restrict__ out) {
    extern __shared__ double2 Zs[];
    const int ring = ring_list[blockIdx.x], sh0 = blockIdx.y * 2 * sch;
    const int n = rt.nphi[ring], start = rt.start[ring], s = rt.shift[ring];
    const double2* tw = rt.tw + rt.twoff[ring];
    // Hermitian-symmetrised spectrum H_a + i H_b, H[k] = ½ Σ_{m≡k} v_m + ½ conj Σ_{m≡-k} v_m, v_m = c_m G_m e^{imφ0}
    for (int x = threadIdx.x; x < n * sch; x += blockDim.x) {
        const int lane = x % sch, k = x / sch;
        const int sa = sh0 + 2 * lane;
        double hr = 0.0, hi = 0.0;  // accumulates (H_a + i H_b)[k]
        if (sa < nrp) {
            for (int pass = 0; pass < 2; ++pass) {
                // pass 0: m ≡ k (direct), pass 1: m ≡ -k (conjugated)
                for (int m = pass ? (n - k) % n : k; m <= lmax; m += n) {
                    const double2 w = tw[(unsigned)(m * s) % (2u * n)];
                    const double cm = (m == 0) ? 0.5 : 1.0;  // ½ (2-δ_m0)
                    const double* g = G + ((size_t)m * nrings + ring) * 2 * nrp + sa;
                    // v = cm * G * e^{+imφ0}
                    const double var = cm * (g[0] * w.x - g[nrp] * w.y), vai = cm * (g[0] * w.y + g[nrp] * w.x);
                    const double vbr = cm * (g[1] * w.x - g[nrp + 1] * w.y), vbi = cm * (g[1] * w.y + g[nrp + 1] * w.x);
                    if (pass == 0) {  // v_a + i v_b
                        hr += var - vbi;
                        hi += vai + vbr;
                    } else {  // conj(v_a) + i conj(v_b)
                        hr += var + vbi;
                        hi += -vai + vbr;
                    }
                }
            }
        }
        Zs[k * sch + lane] = make_double2(hr, hi);
    }
    __syncthreads();
    fft_smem<+1>(Zs, n, log2n, sch, tw);
    for (int x = threadIdx.x; x < n * sch; x += blockDim.x) {
        const int lane = x % sch, j = x / sch;
        const int sa = sh0 + 2 * lane;
        if (sa >= nrp) continue;
        const double2 z = Zs[j * sch + lane];
        const size_t pix = (size_t)start + j;
        double va = 0.0, vb = 0.0;
        if (sa < nr) va = residual ? map[pix * ldw + sa] - z.x : z.x;
        if (sa + 1 < nr) vb = residual ? map[pix * ldw + sa + 1] - z.y : z.y;
        out[pix * nrp + sa] = va;
        out[pix * nrp + sa + 1] = vb;
    }
}

// ---------------------------------------------------------------------------------------------
// Polar-cap rings (nφ = 4i, shifted: φ_j = (j+½)·2π/nφ) with the four-fold mirror symmetry of the ring folded in:
// with q < nφ/4 and the four pixels j1=q, j2=nφ-1-q (φ -> 2π-φ), j3=nφ/2-1-q (φ -> π-φ), j4=nφ/2+q (φ -> π+φ),
//   Re F_m =  Σ_q cos(mφ_q) · (m even ? f1+f2+f3+f4 : f1+f2-f3-f4)
//   Im F_m = -Σ_q sin(mφ_q) · (m even ? f1-f2-f3+f4 : f1-f2+f3-f4)
// so the DFT-as-GEMM runs over nφ/4 points per ring (4x fewer DMMA flops than the unfolded form).
constexpr int kFK = 16;        // folded points (analysis) / m's per parity (synthesis) per smem chunk
constexpr int kLdF = kFK + 4;  // 20 ≡ 4 (mod 16)

// The GEMM kernels below are templated on NI = 8-wide column tiles per warp: the CTA covers BW = 16·NI columns
// (64 for a full shell batch; 32 / 16 when the shells are sharded over GPUs and each rank holds only a few).
template <int NI>
struct ColTile {
    static constexpr int BW = 16 * NI;   // columns per CTA
    static constexpr int LD = BW + 4;    // ≡ 4 (mod 16)
};

// CTA = (cap ring, 32 consecutive m, BW shells).  Row groups of 16: [cos, m even][cos, m odd][-sin, m even][-sin, m odd];
// warp group wm uses its own folded B tile.
template <int NI>
__global__ void __launch_bounds__(kT) cap_analysis_kernel(const double* __restrict__ map, long long ldw, int nr, int nrp,
                                                          RingTabs rt, const int* __restrict__ ring_list, int nrings,
                                                          int lmax, double* __restrict__ F,
                                                          const int* __restrict__ mlim_ring) {
    constexpr int BW = ColTile<NI>::BW, LD = ColTile<NI>::LD;
    __shared__ double As[64 * kLdF];
    __shared__ double Bs[4][kFK * LD];
    // m-chunk index fastest: the CTAs that re-read the same ring of the map are co-scheduled (L2 reuse)
    const int ring = ring_list[blockIdx.y], m0 = blockIdx.x * 32, sh0 = blockIdx.z * BW;
    const int nphi = rt.nphi[ring], start = rt.start[ring];
    // m beyond the ring's cutoff meet only λ_lm < 1e-30 in the Legendre step: their F_m are written as zeros, not computed
    const bool skip = mlim_ring && m0 > mlim_ring[min(ring, nrings - 1 - ring)];
    const int nq = skip ? 0 : (nphi >> 2);
    const double2* tw = rt.tw + rt.twoff[ring];
    const unsigned two_nphi = 2u * nphi;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int wm = warp >> 1, wn = warp & 1;
    double acc[2][NI][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    // software pipeline: twiddle lookups and map loads of chunk q+1 are in flight during the DMMAs of chunk q
    constexpr int NB = (kFK * BW) / kT;
    double2 rw[2];
    double rf[NB][4];
    auto prefetch = [&](int q0) {
        const int kk = tid & 15, q = q0 + kk;
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int m = m0 + (tid >> 4) + 16 * r;
            rw[r] = make_double2(0.0, 0.0);
            if (q < nq && m <= lmax) rw[r] = tw[((unsigned)m * (unsigned)(2 * q + 1)) % two_nphi];
        }
#pragma unroll
        for (int u = 0; u < NB; ++u) {
            const int x = tid + u * kT, k2 = x / BW, c = x % BW, qq = q0 + k2;
            rf[u][0] = rf[u][1] = rf[u][2] = rf[u][3] = 0.0;
            if (qq < nq && sh0 + c < nr) {
                const double* base = map + (size_t)start * ldw + sh0 + c;
                rf[u][0] = base[(size_t)qq * ldw];
                rf[u][1] = base[(size_t)(nphi - 1 - qq) * ldw];
                rf[u][2] = base[(size_t)(nphi / 2 - 1 - qq) * ldw];
                rf[u][3] = base[(size_t)(nphi / 2 + qq) * ldw];
            }
        }
    };
    prefetch(0);
    for (int q0 = 0; q0 < nq; q0 += kFK) {
        {
            const int kk = tid & 15;
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int mm = (tid >> 4) + 16 * r;
                const int row = (mm & 1) * 16 + (mm >> 1);
                As[row * kLdF + kk] = rw[r].x;
                As[(32 + row) * kLdF + kk] = -rw[r].y;
            }
#pragma unroll
            for (int u = 0; u < NB; ++u) {
                const int x = tid + u * kT, k2 = x / BW, c = x % BW;
                const double s = rf[u][0] + rf[u][1], sp = rf[u][2] + rf[u][3];
                const double d = rf[u][0] - rf[u][1], dp = rf[u][2] - rf[u][3];
                Bs[0][k2 * LD + c] = s + sp;
                Bs[1][k2 * LD + c] = s - sp;
                Bs[2][k2 * LD + c] = d - dp;
                Bs[3][k2 * LD + c] = d + dp;
            }
        }
        __syncthreads();
        if (q0 + kFK < nq) prefetch(q0 + kFK);
        warp_gemm_ss<2, NI>(acc, As + wm * 16 * kLdF, kLdF, Bs[wm] + wn * 8 * NI, LD, kFK);
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int rg = i * 8 + g;                    // row within the group of 16
        const int m = m0 + 2 * rg + (wm & 1), comp = wm >> 1;
        if (m > lmax) continue;
        double* dst = F + ((size_t)m * nrings + ring) * 2 * nrp + (size_t)comp * nrp + sh0;
#pragma unroll
        for (int j = 0; j < NI; ++j) {
            const int col = wn * 8 * NI + j * 8 + 2 * t;
            if (sh0 + col < nrp) {
                dst[col] = acc[i][j][0];
                dst[col + 1] = acc[i][j][1];
            }
        }
    }
}

// Synthesis with the same folding: for q < nφ/4
//   Ce = Σ_{m even} c_m ReG_m cos(mφ_q), Co = Σ_{m odd} …, Se = Σ_{m even} c_m ImG_m sin(mφ_q), So = Σ_{m odd} …
//   f1 = Ce+Co-Se-So, f2 = Ce+Co+Se+So, f3 = Ce-Co+Se-So, f4 = Ce-Co-Se+So;   out = residual ? map - f : f
// CTA = (tile = (cap ring, 32 folded points), BW shells); warp group wm computes one of Ce, Co, Se, So.
template <int NI>
__global__ void __launch_bounds__(kT, 2) cap_synthesis_kernel(const double* __restrict__ G, RingTabs rt,
                                                           const int* __restrict__ tile_ring,
                                                           const int* __restrict__ tile_q0, int nrings, int lmax, int nr,
                                                           int nrp, const double* __restrict__ map, long long ldw,
                                                           int residual, double* __restrict__ out) {
    constexpr int BW = ColTile<NI>::BW, LD = ColTile<NI>::LD;
    extern __shared__ double cs_smem[];
    double* As = cs_smem;                    // [4][32][kLdF]   cos even, cos odd, sin even, sin odd
    double* Bs = As + 4 * 32 * kLdF;         // [4][kFK][LD]    Re even, Re odd, Im even, Im odd (times c_m)
    double* Rs = cs_smem;                    // [4][32][LD]     results, aliases the tiles after the K loop
    const int ring = tile_ring[blockIdx.x], q0 = tile_q0[blockIdx.x], sh0 = blockIdx.y * BW;
    const int nphi = rt.nphi[ring], start = rt.start[ring];
    const int nq = nphi >> 2;
    const double2* tw = rt.tw + rt.twoff[ring];
    const unsigned two_nphi = 2u * nphi;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int wm = warp >> 1, wn = warp & 1;
    double acc[4][NI][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    // software pipeline: twiddle lookups and G loads of the next 32 m's are in flight during the DMMAs
    constexpr int NB = (2 * kFK * BW) / kT;
    double2 rw[4];
    double rr[NB], ri[NB];
    auto prefetch = [&](int m0) {
        const int row = tid & 31, q = q0 + row;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int m = m0 + (tid >> 5) + 8 * r;
            rw[r] = make_double2(0.0, 0.0);
            if (q < nq && m <= lmax) rw[r] = tw[((unsigned)m * (unsigned)(2 * q + 1)) % two_nphi];
        }
#pragma unroll
        for (int u = 0; u < NB; ++u) {
            const int x = tid + u * kT, mm = x / BW, c = x % BW, m = m0 + mm;
            rr[u] = ri[u] = 0.0;
            if (m <= lmax && sh0 + c < nrp) {
                const double cm = (m == 0) ? 1.0 : 2.0;
                const double* src = G + ((size_t)m * nrings + ring) * 2 * nrp + sh0 + c;
                rr[u] = cm * src[0];
                ri[u] = cm * src[nrp];
            }
        }
    };
    prefetch(0);
    for (int m0 = 0; m0 <= lmax; m0 += 2 * kFK) {
        {
            const int row = tid & 31;
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int mm = (tid >> 5) + 8 * r;
                const int par = mm & 1, kk = mm >> 1;
                As[(par * 32 + row) * kLdF + kk] = rw[r].x;
                As[((2 + par) * 32 + row) * kLdF + kk] = rw[r].y;
            }
#pragma unroll
            for (int u = 0; u < NB; ++u) {
                const int x = tid + u * kT, mm = x / BW, c = x % BW;
                const int par = mm & 1, kk = mm >> 1;
                Bs[(par * kFK + kk) * LD + c] = rr[u];
                Bs[((2 + par) * kFK + kk) * LD + c] = ri[u];
            }
        }
        __syncthreads();
        if (m0 + 2 * kFK <= lmax) prefetch(m0 + 2 * kFK);
        warp_gemm_ss<4, NI>(acc, As + wm * 32 * kLdF, kLdF, Bs + wm * kFK * LD + wn * 8 * NI, LD, kFK);
        __syncthreads();
    }
    // exchange the four partial results through shared memory
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j) {
            double* dst = Rs + ((size_t)wm * 32 + i * 8 + g) * LD + wn * 8 * NI + j * 8 + 2 * t;
            dst[0] = acc[i][j][0];
            dst[1] = acc[i][j][1];
        }
    __syncthreads();
    for (int x = tid; x < 32 * BW; x += kT) {
        const int row = x / BW, c = x % BW, q = q0 + row, sh = sh0 + c;
        if (q >= nq || sh >= nrp) continue;
        const double ce = Rs[(0 * 32 + row) * LD + c], co = Rs[(1 * 32 + row) * LD + c];
        const double se = Rs[(2 * 32 + row) * LD + c], so = Rs[(3 * 32 + row) * LD + c];
        const size_t p1 = (size_t)start + q, p2 = (size_t)start + nphi - 1 - q;
        const size_t p3 = (size_t)start + nphi / 2 - 1 - q, p4 = (size_t)start + nphi / 2 + q;
        double f1 = 0.0, f2 = 0.0, f3 = 0.0, f4 = 0.0;
        if (sh < nr) {
            f1 = ce + co - se - so;
            f2 = ce + co + se + so;
            f3 = ce - co + se - so;
            f4 = ce - co - se + so;
            if (residual) {
                f1 = map[p1 * ldw + sh] - f1;
                f2 = map[p2 * ldw + sh] - f2;
                f3 = map[p3 * ldw + sh] - f3;
                f4 = map[p4 * ldw + sh] - f4;
            }
        }
        out[p1 * nrp + sh] = f1;
        out[p2 * nrp + sh] = f2;
        out[p3 * nrp + sh] = f3;
        out[p4 * nrp + sh] = f4;
    }
}

template <int NI>
static constexpr int cap_synthesis_smem_bytes() {
    constexpr int tiles = 4 * 32 * kLdF + 4 * kFK * ColTile<NI>::LD, res = 4 * 32 * ColTile<NI>::LD;
    return (tiles > res ? tiles : res) * (int)sizeof(double);
}

// a_lm[c] (+)= w Σ_k λ_lm(θ_k) (F_N ± F_S)[k][c]   CTA = (m, BW columns, 64 l's: 32 of each parity)
template <int NI>
__global__ void __launch_bounds__(kT) legendre_analysis_kernel(const double* __restrict__ F,
                                                               const double* __restrict__ lam, int nrings, int nhalf,
                                                               int kend, int lmax, int nrp, double w, int accumulate,
                                                               const double* __restrict__ add,
                                                               double* __restrict__ alm,
                                                               const int* __restrict__ kbeg_of_m, int dbg) {
    // kend: only the north rings [0, kend) (and their southern mirrors) contribute (kend = nhalf: all rings)
    constexpr int BW = ColTile<NI>::BW, LD = ColTile<NI>::LD;
    extern __shared__ double la_smem[];
    double* As = la_smem;              // [64][kLdA]
    double* Bp = As + 64 * kLdA;       // [32][LD]  F_N + F_S
    double* Bm = Bp + 32 * LD;         // [32][LD]  F_N - F_S
    // l-chunk index fastest: the CTAs that re-read the same F_m slab are co-scheduled and share it through L2
    const int m = blockIdx.z, c0 = blockIdx.y * BW, l0 = m + blockIdx.x * 64;
    if (l0 > lmax) return;
    const int ncol = 2 * nrp;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int wm = warp >> 1, wn = warp & 1;
    const double* Bsel = (wm >> 1) ? Bm : Bp;
    double acc[2][NI][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    const double* Fm = F + (size_t)m * nrings * ncol;

    // software pipeline: the global loads of chunk k+1 are in flight while the DMMAs of chunk k run
    constexpr int NB = (32 * BW) / kT;  // B elements per thread per chunk
    double ra[8], rn[NB], rs[NB];
    size_t arow[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const int row = (tid >> 5) + 8 * q;
        const int l = l0 + ((row < 32) ? 2 * row : 2 * (row - 32) + 1);
        arow[q] = (l <= lmax) ? lm_mmajor(lmax, l, m) * nhalf : (size_t)-1;
    }
    auto prefetch = [&](int k0) {
        const int kk = tid & 31;
        if (dbg & 2) {   // timing aid (SFB_SHT_DBG): no global loads
#pragma unroll
            for (int q = 0; q < 8; ++q) ra[q] = 1.0;
#pragma unroll
            for (int u = 0; u < NB; ++u) rn[u] = rs[u] = 1.0;
            return;
        }
#pragma unroll
        for (int q = 0; q < 8; ++q)
            ra[q] = (arow[q] != (size_t)-1 && k0 + kk < kend) ? lam[arow[q] + k0 + kk] : 0.0;
#pragma unroll
        for (int u = 0; u < NB; ++u) {
            const int x = tid + u * kT, k2 = x / BW, c = x % BW, k = k0 + k2;
            rn[u] = rs[u] = 0.0;
            if (k < kend && c0 + c < ncol) {
                rn[u] = Fm[(size_t)k * ncol + c0 + c];
                if (k != nhalf - 1) rs[u] = Fm[(size_t)(nrings - 1 - k) * ncol + c0 + c];
            }
        }
    };
    // rings [0, kbeg) see only λ_lm < 1e-30 for this m (the cutoff grows towards the equator): start at the first live chunk
    const int kbeg = kbeg_of_m ? min(kbeg_of_m[m], kend) : 0;
    prefetch(kbeg);
    for (int k0 = kbeg; k0 < kend; k0 += 32) {
        {
            const int kk = tid & 31;
#pragma unroll
            for (int q = 0; q < 8; ++q) As[((tid >> 5) + 8 * q) * kLdA + kk] = ra[q];
#pragma unroll
            for (int u = 0; u < NB; ++u) {
                const int x = tid + u * kT, k2 = x / BW, c = x % BW;
                Bp[k2 * LD + c] = rn[u] + rs[u];
                Bm[k2 * LD + c] = rn[u] - rs[u];
            }
        }
        __syncthreads();
        if (k0 + 32 < kend) prefetch(k0 + 32);
        if (!(dbg & 1)) warp_gemm_ss<2, NI>(acc, As + wm * 16 * kLdA, kLdA, Bsel + wn * 8 * NI, LD, 32);
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int row = wm * 16 + i * 8 + g;
        const int l = l0 + ((row < 32) ? 2 * row : 2 * (row - 32) + 1);
        if (l > lmax) continue;
        const size_t off = lm_mmajor(lmax, l, m) * ncol + c0;
        double* dst = alm + off;
#pragma unroll
        for (int j = 0; j < NI; ++j) {
            const int col = wn * 8 * NI + j * 8 + 2 * t;
            if (c0 + col < ncol) {
                double v0 = w * acc[i][j][0], v1 = w * acc[i][j][1];
                if (add) {
                    v0 += add[off + col];
                    v1 += add[off + col + 1];
                }
                if (accumulate) {
                    v0 += dst[col];
                    v1 += dst[col + 1];
                }
                dst[col] = v0;
                dst[col + 1] = v1;
            }
        }
    }
}

// G_m(ring)[c] = Σ_l λ_lm(θ) a_lm[c]: north = E + O, south = E - O    CTA = (m, 64 north rings, BW columns)
template <int NI>
__global__ void __launch_bounds__(kT, 2) legendre_synthesis_kernel(const double* __restrict__ alm,
                                                                const double* __restrict__ lam, int nrings, int nhalf,
                                                                int kend, int lmax, int nrp, double* __restrict__ G,
                                                                double* __restrict__ F2,
                                                                const int* __restrict__ nphi_tab,
                                                                const int* __restrict__ mlim_tile) {
    // kend: only the north rings [0, kend) and their mirrors are synthesised (the grid covers ceil(kend/64) tiles)
    // F2 (optional): rings with nφ > 2 lmax have no aliases, their re-analysed Fourier coefficients are F'_m = nφ G_m
    // (imaginary part of m = 0 dropped; see ring_alias_kernel), so they are written there directly and the alias pass
    // only visits the short polar rings.
    constexpr int BW = ColTile<NI>::BW, LD = ColTile<NI>::LD;
    __shared__ double Ls[32 * kLdB];  // [l (16 even-parity, 16 odd-parity)][ring]
    __shared__ double Bs[32 * LD];
    const int m = blockIdx.x, k0 = blockIdx.y * 64, c0 = blockIdx.z * BW;
    // every ring of this tile is beyond the cutoff of m: its G_m would be < 1e-30 |alm|; the consumers (alias pass with the
    // same cutoff, Legendre analysis) never use it.  Only passed in the ring-space Jacobi passes (alm2map needs every G).
    if (mlim_tile && m > mlim_tile[blockIdx.y]) return;
    const int ncol = 2 * nrp;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int wm = warp >> 1, wn = warp & 1;
    double accE[2][NI][2], accO[2][NI][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j) accE[i][j][0] = accE[i][j][1] = accO[i][j][0] = accO[i][j][1] = 0.0;

    // software pipeline: loads of the next 32 l's are in flight while the DMMAs of the current chunk run
    constexpr int NB = (32 * BW) / kT;
    double rl[8], rb[NB];
    auto prefetch = [&](int l0) {
        const int x = tid & 63;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int k2 = (tid >> 6) + 4 * q;
            const int l = l0 + ((k2 < 16) ? 2 * k2 : 2 * (k2 - 16) + 1);
            rl[q] = (l <= lmax && k0 + x < kend) ? lam[lm_mmajor(lmax, l, m) * nhalf + k0 + x] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < NB; ++u) {
            const int y = tid + u * kT, k2 = y / BW, c = y % BW;
            const int l = l0 + ((k2 < 16) ? 2 * k2 : 2 * (k2 - 16) + 1);
            rb[u] = (l <= lmax && c0 + c < ncol) ? alm[lm_mmajor(lmax, l, m) * ncol + c0 + c] : 0.0;
        }
    };
    prefetch(m);
    for (int l0 = m; l0 <= lmax; l0 += 32) {
        {
            const int x = tid & 63;
#pragma unroll
            for (int q = 0; q < 8; ++q) Ls[((tid >> 6) + 4 * q) * kLdB + x] = rl[q];
#pragma unroll
            for (int u = 0; u < NB; ++u) {
                const int y = tid + u * kT;
                Bs[(y / BW) * LD + (y % BW)] = rb[u];
            }
        }
        __syncthreads();
        if (l0 + 32 <= lmax) prefetch(l0 + 32);
        warp_gemm_ts<2, NI>(accE, Ls + wm * 16, kLdB, Bs + wn * 8 * NI, LD, 16);
        warp_gemm_ts<2, NI>(accO, Ls + 16 * kLdB + wm * 16, kLdB, Bs + 16 * LD + wn * 8 * NI, LD, 16);
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int k = k0 + wm * 16 + i * 8 + g;
        if (k >= kend) continue;
        const int nph = F2 ? nphi_tab[k] : 0;
        const bool direct = nph > 2 * lmax;
        const double sc = direct ? (double)nph : 1.0;
        double* Gm = (direct ? F2 : G) + (size_t)m * nrings * ncol;
        double* dn = Gm + (size_t)k * ncol + c0;
        double* ds = Gm + (size_t)(nrings - 1 - k) * ncol + c0;
#pragma unroll
        for (int j = 0; j < NI; ++j) {
            const int col = wn * 8 * NI + j * 8 + 2 * t;
            if (c0 + col < ncol) {
                const double z = (direct && m == 0 && c0 + col >= nrp) ? 0.0 : sc;  // nrp is even: col, col+1 same plane
                dn[col] = z * (accE[i][j][0] + accO[i][j][0]);
                dn[col + 1] = z * (accE[i][j][1] + accO[i][j][1]);
                if (k != nhalf - 1) {
                    ds[col] = z * (accE[i][j][0] - accO[i][j][0]);
                    ds[col + 1] = z * (accE[i][j][1] - accO[i][j][1]);
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Alias-free rings through a precomputed Gram matrix.  On a ring with nφ > 2 lmax the re-analysed coefficients are
// F'_m = nφ G_m, so the part of a Jacobi pass that runs over those rings,  alm -> w Λ_direct nφ Λ_directᵀ alm,  is the
// data-independent operator (block diagonal in m; even and odd l-m decouple because λ_lm(π-θ) = (-1)^{l+m} λ_lm(θ))
//   K_m[l][l'] = w Σ_{north rings k >= kpolar} mult_k nφ_k λ_lm(θ_k) λ_l'm(θ_k),   mult = 2 (1 on the equator).
// A pass applies K_m (2 L² flops per column and m, L = number of l of one parity) instead of going through the rings
// twice (4 L · nrings): ≈ 8x fewer flops for these rings at cfg4.  Only the polar rings [0, kpolar) keep
// synthesis -> alias -> analysis.  (Reduction checked on CPU: tests/test_oracle_sht.py::test_belt_rings_reduce_to_gram_matrices.)
//
// Storage: block (m, par) at gram_off[2m+par], row-major [Lp][ldk], ldk = Lp rounded up to 32, zero padded columns.

// K tile 32x32 per CTA, plain FMAs (runs once per plan).  grid = (col tiles, row tiles, 2(lmax+1))
__global__ void __launch_bounds__(256) gram_build_kernel(const double* __restrict__ lam, const int* __restrict__ nphi_tab,
                                                         const long long* __restrict__ gram_off, int lmax, int nhalf,
                                                         int kpolar, double w, double* __restrict__ K) {
    const int m = blockIdx.z >> 1, par = blockIdx.z & 1;
    const int Lp = (m + par <= lmax) ? (lmax - m - par) / 2 + 1 : 0;
    const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
    if (i0 >= Lp || j0 >= Lp) return;
    const int ldk = (Lp + 31) & ~31;
    __shared__ double Ai[32][33], Aj[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (int k0 = kpolar; k0 < nhalf; k0 += 32) {
        for (int r = ty; r < 32; r += 8) {
            const int k = k0 + tx;
            const int li = m + par + 2 * (i0 + r), lj = m + par + 2 * (j0 + r);
            double sc = 0.0;
            if (k < nhalf) sc = w * (double)nphi_tab[k] * ((k == nhalf - 1) ? 1.0 : 2.0);
            Ai[r][tx] = (k < nhalf && li <= lmax) ? sc * lam[lm_mmajor(lmax, li, m) * nhalf + k] : 0.0;
            Aj[r][tx] = (k < nhalf && lj <= lmax) ? lam[lm_mmajor(lmax, lj, m) * nhalf + k] : 0.0;
        }
        __syncthreads();
#pragma unroll 8
        for (int kk = 0; kk < 32; ++kk) {
            const double b = Aj[tx][kk];
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[q] = fma(Ai[ty + 8 * q][kk], b, acc[q]);
        }
        __syncthreads();
    }
    double* dst = K + gram_off[blockIdx.z];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int i = i0 + ty + 8 * q, j = j0 + tx;
        if (i < Lp && j < ldk) dst[(size_t)i * ldk + j] = (j < Lp) ? acc[q] : 0.0;
    }
}

// out[l][c] = add[l][c] - Σ_l' K_m[l][l'] x[l'][c]      CTA = ((m, parity), 64 l's of that parity, BW columns)
template <int NI>
__global__ void __launch_bounds__(kT) gram_apply_kernel(const double* __restrict__ K, const long long* __restrict__ gram_off,
                                                        const double* __restrict__ x, const double* __restrict__ add,
                                                        int lmax, int nrp, double* __restrict__ out) {
    constexpr int BW = ColTile<NI>::BW, LD = ColTile<NI>::LD;
    __shared__ double As[64 * kLdA];
    __shared__ double Bs[32 * LD];
    const int m = blockIdx.z >> 1, par = blockIdx.z & 1;
    const int Lp = (m + par <= lmax) ? (lmax - m - par) / 2 + 1 : 0;
    const int r0 = blockIdx.x * 64, c0 = blockIdx.y * BW;
    if (r0 >= Lp) return;
    const int ldk = (Lp + 31) & ~31, ncol = 2 * nrp;
    const double* Km = K + gram_off[blockIdx.z];
    const size_t lm0 = lm_mmajor(lmax, m + par, m);            // alm row of the first l of this parity; rows step by 2
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int wm = warp >> 1, wn = warp & 1;
    double acc[2][NI][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    constexpr int NB = (32 * BW) / kT;
    double ra[8], rb[NB];
    auto prefetch = [&](int k0) {
        const int kk = tid & 31;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int row = r0 + (tid >> 5) + 8 * q;
            ra[q] = (row < Lp) ? Km[(size_t)row * ldk + k0 + kk] : 0.0;      // k0 + kk < ldk always (zero padded)
        }
#pragma unroll
        for (int u = 0; u < NB; ++u) {
            const int y = tid + u * kT, k2 = y / BW, c = y % BW, j = k0 + k2;
            rb[u] = (j < Lp && c0 + c < ncol) ? x[(lm0 + 2 * (size_t)j) * ncol + c0 + c] : 0.0;
        }
    };
    prefetch(0);
    for (int k0 = 0; k0 < Lp; k0 += 32) {
        {
            const int kk = tid & 31;
#pragma unroll
            for (int q = 0; q < 8; ++q) As[((tid >> 5) + 8 * q) * kLdA + kk] = ra[q];
#pragma unroll
            for (int u = 0; u < NB; ++u) {
                const int y = tid + u * kT;
                Bs[(y / BW) * LD + (y % BW)] = rb[u];
            }
        }
        __syncthreads();
        if (k0 + 32 < Lp) prefetch(k0 + 32);
        warp_gemm_ss<2, NI>(acc, As + wm * 16 * kLdA, kLdA, Bs + wn * 8 * NI, LD, 32);
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int row = r0 + wm * 16 + i * 8 + g;
        if (row >= Lp) continue;
        const size_t off = (lm0 + 2 * (size_t)row) * ncol + c0;
#pragma unroll
        for (int j = 0; j < NI; ++j) {
            const int col = wn * 8 * NI + j * 8 + 2 * t;
            if (c0 + col < ncol) {
                out[off + col] = add[off + col] - acc[i][j][0];
                out[off + col + 1] = add[off + col + 1] - acc[i][j][1];
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Ring-space form of the Jacobi refinement.  Re-analysing a synthesised ring needs no pixels: with
// f_j = Σ_{m'} c_{m'} Re(G_{m'} e^{im'φ_j}) and Σ_j e^{ikφ_j} = nφ e^{ikφ0} [k ≡ 0 mod nφ],
//   F'_m = Σ_j f_j e^{-imφ_j}
//        = nφ Σ_{m'≡m} (c_{m'}/2) G_{m'} σ^{(m'-m)/nφ} + nφ Σ_{m'≡-m} (c_{m'}/2) conj(G_{m'}) σ^{(m'+m)/nφ},
// m' in [0, lmax], congruences mod nφ, σ = -1 on shifted rings (φ0 = π/nφ) and +1 otherwise.  For nφ > 2 lmax this is
// F'_m = nφ G_m; short polar rings pick up their exact aliases.  So map - S(alm) is never formed in the iterations:
//   alm <- alm + A f - LegendreAnalysis(F'(LegendreSynthesis(alm))).
// thread = (ring, m, column c of the re plane); G and F2 are [m][ring][re/im][nrp].
// Class-sum form (identity checked on CPU: tests/test_oracle_sht.py::test_alias_operator_class_sum_form).
// F'_m depends on m only through ρ = m mod nφ and q = m div nφ: with the alternating class sums
//   Q[ρ] = Σ_j σ^j c_{ρ+j nφ} G_{ρ+j nφ}      (c_0 = 1/2, c_{m'} = 1 otherwise; σ = -1 on shifted rings)
//   F'_m = nφ σ^q ( Q[ρ] + σ^{[ρ≠0]} conj Q[(nφ-ρ) mod nφ] ).
// CTA = (aliased ring, 16 shells): one pass over m' builds Q in shared memory (every G element is read exactly once,
// the per-(ring, m) kernel below reads it once per aliased m: 140 times on ring 1), one pass over m writes F'.
constexpr int kAliasCW = 16;
__global__ void __launch_bounds__(256) ring_alias_smem_kernel(const double* __restrict__ G, double* __restrict__ F2,
                                                              const int* __restrict__ ring_list,
                                                              const int* __restrict__ nphi_tab,
                                                              const int* __restrict__ shift_tab, int nrings, int lmax,
                                                              int nrp, const int* __restrict__ mlim_ring) {
    extern __shared__ double Qs[];   // [ρ < min(nφ, lmax+1)][re/im][kAliasCW]
    constexpr int CW = kAliasCW;
    const int ring = ring_list[blockIdx.x], c0 = blockIdx.y * CW;
    const int n = nphi_tab[ring];
    const double sig = shift_tab[ring] ? -1.0 : 1.0;
    const size_t stride_m = (size_t)nrings * 2 * nrp;
    const double* g = G + (size_t)ring * 2 * nrp + c0;
    const int nrho = min(n, lmax + 1);
    for (int x = threadIdx.x; x < nrho * 2 * CW; x += blockDim.x) {
        const int c = x % CW, comp = (x / CW) & 1, rho = x / (2 * CW);
        double q = 0.0;
        if (c0 + c < nrp) {
            double s = (rho == 0) ? 0.5 : 1.0;
            const int mcap = mlim_ring ? min(lmax, mlim_ring[min(ring, nrings - 1 - ring)]) : lmax;   // G beyond: not synthesised
            for (int mp = rho; mp <= mcap; mp += n) {
                q = fma(s, g[(size_t)mp * stride_m + (size_t)comp * nrp + c], q);
                s = (mp == 0) ? sig : s * sig;           // c_0 = 1/2 applies to m' = 0 only
            }
        }
        Qs[x] = q;
    }
    __syncthreads();
    double* f = F2 + (size_t)ring * 2 * nrp + c0;
    for (int x = threadIdx.x; x < (lmax + 1) * CW; x += blockDim.x) {
        const int c = x % CW, m = x / CW;
        if (c0 + c >= nrp) continue;
        const int rho = m % n, q = m / n, rho2 = (n - rho) % n;
        const double sq = (q & 1) ? sig : 1.0, s2 = (rho != 0) ? sig : 1.0;
        const double re1 = Qs[(rho * 2) * CW + c], im1 = Qs[(rho * 2 + 1) * CW + c];
        double re2 = 0.0, im2 = 0.0;
        if (rho2 < nrho) {
            re2 = Qs[(rho2 * 2) * CW + c];
            im2 = Qs[(rho2 * 2 + 1) * CW + c];
        }
        f[(size_t)m * stride_m + c] = n * sq * (re1 + s2 * re2);
        f[(size_t)m * stride_m + nrp + c] = n * sq * (im1 - s2 * im2);
    }
}

__global__ void ring_alias_kernel(const double* __restrict__ G, double* __restrict__ F2, const int* __restrict__ nphi_tab,
                                  const int* __restrict__ shift_tab, int nrings, int lmax, int nrp) {
    const int ring = blockIdx.x, m = blockIdx.y;
    const int n = nphi_tab[ring];
    if (n > 2 * lmax) return;  // alias free: F' = nφ G was written by legendre_synthesis_kernel
    const double sig = shift_tab[ring] ? -1.0 : 1.0;
    const size_t stride_m = (size_t)nrings * 2 * nrp;
    const double* g = G + (size_t)ring * 2 * nrp;
    double* f = F2 + (size_t)m * stride_m + (size_t)ring * 2 * nrp;
    for (int c = threadIdx.x; c < nrp; c += blockDim.x) {
        double re = 0.0, im = 0.0;
        // m' ≡ m  (mod n)
        for (int mp = m % n; mp <= lmax; mp += n) {
            const int tt = (mp - m) / n;  // may be negative
            const double s = ((tt & 1) ? sig : 1.0) * ((mp == 0) ? 0.5 : 1.0);
            re += s * g[(size_t)mp * stride_m + c];
            im += s * g[(size_t)mp * stride_m + nrp + c];
        }
        // m' ≡ -m (mod n)
        for (int mp = (n - m % n) % n; mp <= lmax; mp += n) {
            const int tt = (mp + m) / n;
            const double s = ((tt & 1) ? sig : 1.0) * ((mp == 0) ? 0.5 : 1.0);
            re += s * g[(size_t)mp * stride_m + c];
            im -= s * g[(size_t)mp * stride_m + nrp + c];
        }
        f[c] = n * re;
        f[nrp + c] = n * im;
    }
}

// column-tile width class for `cols` live columns
static inline int pick_ni(int cols) { return cols > 32 ? 4 : (cols > 16 ? 2 : 1); }

// planar [lm m-major][comp][nrp] -> interleaved complex, column-major nr x lmsize, requested column order
__global__ void alm_to_complex_kernel(const double* __restrict__ alm, int lmax, int nr, int nrp, int layout,
                                      double* __restrict__ out) {
    const int l = blockIdx.x, m = blockIdx.y;
    if (m > l) return;
    const size_t src = lm_mmajor(lmax, l, m);
    const size_t dst = layout ? ((size_t)l * (l + 1) / 2 + m) : src;
    for (int r = threadIdx.x; r < nr; r += blockDim.x) {
        out[2 * (dst * nr + r)] = alm[src * 2 * nrp + r];
        out[2 * (dst * nr + r) + 1] = alm[src * 2 * nrp + nrp + r];
    }
}

// =============================================================================================
// host side

int sht_plan_create(ShtPlan** out, int64_t nside_in, int64_t nside_out, int64_t lmax, int64_t nr) {
    SFB_REQUIRE(out, "sht_plan_create: null pointer");
    auto pow2 = [](int64_t v) { return v >= 1 && (v & (v - 1)) == 0; };
    SFB_REQUIRE(nside_out >= 1 && nside_in >= 1 && nside_out <= 8192 && nside_in <= 8192,
                "sht_plan_create: nside out of range");
    SFB_REQUIRE(nside_in == nside_out || (pow2(nside_in) && pow2(nside_out)),
                "udgrade: nside must be a power of two when the resolution changes");
    SFB_REQUIRE(lmax >= 0 && nr >= 1, "sht_plan_create: bad sizes");
    if (lmax > 4 * nside_out) {  // src/healpix_helpers.jl:60-63
        set_error("lmax > 4*nside is a poor choice (lmax=" + std::to_string(lmax) + ", 4*nside=" +
                  std::to_string(4 * nside_out) + ")");
        return 3;
    }
    SFB_REQUIRE(lmax <= 1500, "sht_plan_create: lmax > 1500 needs a scaled Legendre recurrence (not in this build)");
    auto* p = new ShtPlan();
    p->nside_in = (int)nside_in;
    p->nside = (int)nside_out;
    p->lmax = (int)lmax;
    p->nr = (int)nr;
    p->nrp = (int)round_up(nr, 8);
    p->npix_in = 12 * nside_in * nside_in;
    p->npix = 12 * nside_out * nside_out;
    p->nrings = 4 * p->nside - 1;
    p->nhalf = 2 * p->nside;
    p->lmsize = (size_t)(lmax + 1) * (lmax + 2) / 2;

    const int ns = p->nside;
    std::vector<int> nphi(p->nrings), start(p->nrings), shift(p->nrings), twoff(p->nrings), tile_ring, tile_j0;
    std::vector<int> gemm_rings, fft_rings, cap_rings, ctile_ring, ctile_q0;
    // belt rings (nφ = 4 nside) go through the shared-memory FFT when nside is a power of two
    p->use_fft = pow2(nside_out) && (4 * nside_out >= 8) && (4 * nside_out <= 4096);  // <= 16 values per thread
    p->log2n = 0;
    while ((1 << p->log2n) < 4 * p->nside) p->log2n++;
    const int64_t ncap = 2LL * ns * (ns - 1);
    for (int idx = 0; idx < p->nrings; ++idx) {
        const int i = idx + 1, north = (i <= 2 * ns) ? i : 4 * ns - i;
        int64_t st;
        int np, sh, slot;
        if (north < ns) {
            np = 4 * north;
            st = 2LL * north * (north - 1);
            sh = 1;
            slot = north;
        } else {
            np = 4 * ns;
            st = ncap + (int64_t)(north - ns) * 4 * ns;
            sh = (((north - ns) & 1) == 0) ? 1 : 0;
            slot = ns;
        }
        if (i > 2 * ns) st = p->npix - st - np;
        nphi[idx] = np;
        start[idx] = (int)st;
        shift[idx] = sh;
        twoff[idx] = 4 * slot * (slot - 1);
        const bool fft = p->use_fft && north >= ns;
        const bool cap = north < ns;  // shifted ring with nφ = 4i: four-fold folded DFT-as-GEMM
        (fft ? fft_rings : (cap ? cap_rings : gemm_rings)).push_back(idx);
        if (cap)
            for (int q0 = 0; q0 < np / 4; q0 += 32) {
                ctile_ring.push_back(idx);
                ctile_q0.push_back(q0);
            }
        else if (!fft)
            for (int j0 = 0; j0 < np; j0 += 64) {
                tile_ring.push_back(idx);
                tile_j0.push_back(j0);
            }
    }
    p->n_cap_rings = (int)cap_rings.size();
    p->n_ctiles = (int)ctile_ring.size();
    p->ntiles = (int)tile_ring.size();
    p->n_gemm_rings = (int)gemm_rings.size();
    p->n_fft_rings = (int)fft_rings.size();
    {
        // complex lanes per CTA: n * sch * 16 B <= 64 KB, a power of two dividing nrp/2
        const int fft_budget = getenv("SFB_FFT_ELEMS") ? atoi(getenv("SFB_FFT_ELEMS")) : 4096;  // complex values per CTA
        int bound = std::max(1, fft_budget / (4 * p->nside));
        bound = std::min(bound, 16);
        int sch = 1;
        while (sch * 2 <= bound && (p->nrp / 2) % (sch * 2) == 0) sch *= 2;
        p->fft_sch = sch;
    }
    auto up = [&](DevBuf<int>& d, const std::vector<int>& h) -> int {
        SFB_TRY(d.alloc(h.size()));
        SFB_CUDA_OK(cudaMemcpy(d.p, h.data(), h.size() * sizeof(int), cudaMemcpyHostToDevice));
        return 0;
    };
    int rc = 0;
    rc = rc ? rc : up(p->d_nphi, nphi);
    rc = rc ? rc : up(p->d_start, start);
    rc = rc ? rc : up(p->d_shift, shift);
    rc = rc ? rc : up(p->d_twoff, twoff);
    rc = rc ? rc : up(p->d_tile_ring, tile_ring);
    rc = rc ? rc : up(p->d_gemm_rings, gemm_rings);
    rc = rc ? rc : up(p->d_fft_rings, fft_rings);
    rc = rc ? rc : up(p->d_cap_rings, cap_rings);
    rc = rc ? rc : up(p->d_ctile_ring, ctile_ring);
    rc = rc ? rc : up(p->d_ctile_q0, ctile_q0);
    rc = rc ? rc : up(p->d_tile_j0, tile_j0);
    rc = rc ? rc : p->d_tw.alloc((size_t)4 * ns * (ns + 1));
    rc = rc ? rc : p->d_lam.alloc(p->lmsize * p->nhalf);
    rc = rc ? rc : p->d_FG.alloc((size_t)(lmax + 1) * p->nrings * 2 * p->nrp);
    if (!rc && nside_in != nside_out) rc = p->d_map.alloc((size_t)p->npix * p->nrp);
    if (rc) {
        delete p;
        return rc;
    }
    twiddle_table_kernel<<<ns, 256>>>(p->d_tw.p, ns);
    lambda_table_kernel<<<dim3((unsigned)ceil_div(p->nhalf, 128), (unsigned)(lmax + 1)), 128>>>(p->d_lam.p, ns, (int)lmax,
                                                                                              p->nhalf);
    if (!getenv("SFB_SHT_NO_MLIM")) {
        // m-cutoff tables from the λ table itself: mlim_ring[k] = largest m with max_l |λ_lm(θ_k)| >= 1e-30
        DevBuf<double> d_mx;
        rc = d_mx.alloc((size_t)(lmax + 1) * p->nhalf);
        if (!rc) {
            lambda_absmax_kernel<<<dim3((unsigned)ceil_div(p->nhalf, 128), (unsigned)(lmax + 1)), 128>>>(p->d_lam.p, (int)lmax,
                                                                                                         p->nhalf, d_mx.p);
            std::vector<double> mx((size_t)(lmax + 1) * p->nhalf);
            if (cudaMemcpy(mx.data(), d_mx.p, mx.size() * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess) rc = 1;
            if (!rc) {
                std::vector<int> mlim(p->nhalf, 0);
                for (int k = 0; k < p->nhalf; ++k)
                    for (int m = 0; m <= lmax; ++m)
                        if (mx[(size_t)m * p->nhalf + k] >= 1e-30) mlim[k] = m;
                for (int k = 1; k < p->nhalf; ++k) mlim[k] = std::max(mlim[k], mlim[k - 1]);   // monotone towards the equator
                std::vector<int> tile((p->nhalf + 63) / 64, 0), kbeg(lmax + 1, 0);
                for (int k = 0; k < p->nhalf; ++k) tile[k / 64] = std::max(tile[k / 64], mlim[k]);
                for (int m = 0; m <= lmax; ++m) {
                    int k = 0;
                    while (k < p->nhalf && mlim[k] < m) ++k;       // first ring that sees m
                    kbeg[m] = (k / 32) * 32;                        // chunk of 32 rings it lives in
                }
                rc = rc ? rc : up(p->d_mlim_ring, mlim);
                rc = rc ? rc : up(p->d_mlim_tile, tile);
                rc = rc ? rc : up(p->d_kbeg_of_m, kbeg);
            }
        }
        if (rc) {
            set_error("sht_plan_create: m-cutoff tables");
            delete p;
            return rc;
        }
    }
    // skipped (ring, m) pairs leave their slots untouched: the buffers must hold finite numbers from the start
    if (cudaMemset(p->d_FG.p, 0, p->d_FG.n * sizeof(double)) != cudaSuccess) {
        set_error("sht_plan_create: cudaMemset");
        delete p;
        return 1;
    }
    {
        // rings [0, kpolar) keep synthesis -> alias -> analysis in a Jacobi pass; the alias-free rings beyond go through
        // the Gram matrices.  kpolar = the aliased north rings (nφ <= 2 lmax) rounded up to the synthesis tile of 64.
        int kalias = 0;
        for (int k = 0; k < p->nhalf; ++k)
            if (nphi[k] <= 2 * lmax) kalias = k + 1;
        std::vector<int> alias_rings;     // both hemispheres, shortest rings last (they do the most work per CTA: first)
        for (int idx = 0; idx < p->nrings; ++idx)
            if (nphi[idx] <= 2 * lmax) alias_rings.push_back(idx);
        std::sort(alias_rings.begin(), alias_rings.end(), [&](int a, int b) { return nphi[a] < nphi[b]; });
        p->n_alias_rings = (int)alias_rings.size();
        rc = up(p->d_alias_rings, alias_rings);
        if (rc) {
            delete p;
            return rc;
        }
        p->kpolar = (int)std::min<int64_t>(p->nhalf, round_up(kalias, 64));
        if (getenv("SFB_SHT_NO_GRAM")) p->kpolar = p->nhalf;
        if (p->kpolar < p->nhalf) {
            std::vector<long long> off(2 * (lmax + 1) + 1, 0);
            for (int m = 0; m <= lmax; ++m)
                for (int par = 0; par < 2; ++par) {
                    const long long Lp = (m + par <= lmax) ? (lmax - m - par) / 2 + 1 : 0;
                    off[2 * m + par + 1] = off[2 * m + par] + Lp * round_up(Lp, 32);
                }
            rc = p->d_gram_off.alloc(off.size());
            if (!rc && cudaMemcpy(p->d_gram_off.p, off.data(), off.size() * sizeof(long long), cudaMemcpyHostToDevice) !=
                           cudaSuccess)
                rc = 1;
            rc = rc ? rc : p->d_gram.alloc((size_t)std::max<long long>(1, off.back()));
            if (rc) {
                set_error("sht_plan_create: Gram-matrix tables");
                delete p;
                return rc;
            }
            const unsigned tiles = (unsigned)ceil_div(lmax / 2 + 1, 32);
            gram_build_kernel<<<dim3(tiles, tiles, 2 * (unsigned)(lmax + 1)), 256>>>(
                p->d_lam.p, p->d_nphi.p, p->d_gram_off.p, (int)lmax, p->nhalf, p->kpolar,
                4.0 * 3.14159265358979323846 / (double)p->npix, p->d_gram.p);
        }
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        set_error(std::string("sht_plan_create setup kernels: ") + cudaGetErrorString(e));
        delete p;
        return 1;
    }
    *out = p;
    return 0;
}

void sht_plan_destroy(ShtPlan* p) {
    if (p && p->pending) {
        cudaEventDestroy(p->tev0);
        cudaEventDestroy(p->tev1);
    }
    if (p && p->side) {
        cudaStreamDestroy(p->side);
        cudaEventDestroy(p->fork_ev);
        cudaEventDestroy(p->join_ev);
    }
    delete p;
}

// fork: work enqueued on p->side after this call starts once everything enqueued on `st` so far has finished;
// join: `st` continues once the side stream has drained.  Returns false (and does nothing) in serial mode.
static bool sht_fork(ShtPlan* p, cudaStream_t st) {
    if (getenv("SFB_SHT_SERIAL")) return false;
    if (!p->side) {
        if (cudaStreamCreateWithFlags(&p->side, cudaStreamNonBlocking) != cudaSuccess) return false;
        cudaEventCreateWithFlags(&p->fork_ev, cudaEventDisableTiming);
        cudaEventCreateWithFlags(&p->join_ev, cudaEventDisableTiming);
    }
    if (cudaEventRecord(p->fork_ev, st) != cudaSuccess) return false;
    return cudaStreamWaitEvent(p->side, p->fork_ev, 0) == cudaSuccess;
}
static int sht_join(ShtPlan* p, cudaStream_t st) {
    SFB_CUDA_OK(cudaEventRecord(p->join_ev, p->side));
    SFB_CUDA_OK(cudaStreamWaitEvent(st, p->join_ev, 0));
    return 0;
}

// map -> ring-Fourier coefficients F (d_FG)
static int run_ring_analysis(ShtPlan* p, const double* map, int64_t ldw, cudaStream_t st) {
    const int lmax = p->lmax, nrp = p->nrp;
    RingTabs rt{p->d_nphi.p, p->d_start.p, p->d_shift.p, p->d_twoff.p, p->d_tw.p};
    // the belt FFT (latency/HBM-bound) and the cap DFT (DMMA) touch disjoint rings: side by side on two streams
    const bool forked = p->n_fft_rings > 0 && (p->n_cap_rings > 0 || p->n_gemm_rings > 0) && sht_fork(p, st);
    if (p->n_gemm_rings > 0) {
        dim3 g1(p->n_gemm_rings, (unsigned)ceil_div(lmax + 1, 32), (unsigned)ceil_div(nrp, 64));
        ring_analysis_kernel<<<g1, kT, 0, st>>>(map, ldw, p->nr, nrp, rt, p->d_gemm_rings.p, p->nrings, lmax, p->d_FG.p);
        SFB_CUDA_OK(cudaGetLastError());
        p->launches++;
    }
    if (p->n_cap_rings > 0) {
        const int ni = pick_ni(nrp);
        dim3 gc((unsigned)ceil_div(lmax + 1, 32), p->n_cap_rings, (unsigned)ceil_div(nrp, 16 * ni));
        if (ni == 4)
            cap_analysis_kernel<4><<<gc, kT, 0, st>>>(map, ldw, p->nr, nrp, rt, p->d_cap_rings.p, p->nrings, lmax, p->d_FG.p,
                                                      p->d_mlim_ring.p);
        else if (ni == 2)
            cap_analysis_kernel<2><<<gc, kT, 0, st>>>(map, ldw, p->nr, nrp, rt, p->d_cap_rings.p, p->nrings, lmax, p->d_FG.p,
                                                      p->d_mlim_ring.p);
        else
            cap_analysis_kernel<1><<<gc, kT, 0, st>>>(map, ldw, p->nr, nrp, rt, p->d_cap_rings.p, p->nrings, lmax, p->d_FG.p,
                                                      p->d_mlim_ring.p);
        SFB_CUDA_OK(cudaGetLastError());
        p->launches++;
    }
    if (p->n_fft_rings > 0) {
        const int n = 4 * p->nside, sch = p->fft_sch;
        const int smem = n * sch * (int)sizeof(double2);
        SFB_CUDA_OK(cudaFuncSetAttribute(belt_analysis_fft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        dim3 gf(p->n_fft_rings, (unsigned)ceil_div(nrp, 2 * sch));
        belt_analysis_fft_kernel<<<gf, kT, smem, forked ? p->side : st>>>(map, ldw, p->nr, nrp, rt, p->d_fft_rings.p,
                                                                          p->nrings, lmax, p->log2n, sch, p->d_FG.p);
        SFB_CUDA_OK(cudaGetLastError());
        p->launches++;
    }
    if (forked) SFB_TRY(sht_join(p, st));
    return 0;
}

// F (ring-Fourier coefficients, in `Fsrc`) -> alm:  alm = [accumulate ? alm : 0] + [add ? add : 0] + w Λ F
static int run_legendre_analysis(ShtPlan* p, const double* Fsrc, double w, int accumulate, const double* add,
                                 double* d_alm, cudaStream_t st, int kend = -1) {
    if (kend < 0) kend = p->nhalf;
    const int lmax = p->lmax, nrp = p->nrp;
    const int nil = pick_ni(2 * nrp);
    dim3 g2((unsigned)ceil_div(lmax + 1, 64), (unsigned)ceil_div(2 * nrp, 16 * nil), lmax + 1);
    const int la_smem_bytes = (64 * kLdA + 2 * 32 * (16 * nil + 4)) * (int)sizeof(double);
    const int sht_dbg = getenv("SFB_SHT_DBG") ? atoi(getenv("SFB_SHT_DBG")) : 0;   // timing aid, results wrong when != 0
#define SFB_LAUNCH_LA(NI_)                                                                                            \
    do {                                                                                                              \
        SFB_CUDA_OK(cudaFuncSetAttribute(legendre_analysis_kernel<NI_>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                         la_smem_bytes));                                                             \
        legendre_analysis_kernel<NI_><<<g2, kT, la_smem_bytes, st>>>(Fsrc, p->d_lam.p, p->nrings, p->nhalf, kend, lmax,  \
                                                                     nrp, w, accumulate, add, d_alm,                 \
                                                                     p->d_kbeg_of_m.p, sht_dbg);                     \
    } while (0)
    if (nil == 4)
        SFB_LAUNCH_LA(4);
    else if (nil == 2)
        SFB_LAUNCH_LA(2);
    else
        SFB_LAUNCH_LA(1);
#undef SFB_LAUNCH_LA
    SFB_CUDA_OK(cudaGetLastError());
    p->launches += 1;
    return 0;
}

static int run_analysis(ShtPlan* p, const double* map, int64_t ldw, int accumulate, double* d_alm, cudaStream_t st) {
    SFB_TRY(run_ring_analysis(p, map, ldw, st));
    return run_legendre_analysis(p, p->d_FG.p, 4.0 * 3.14159265358979323846 / (double)p->npix, accumulate, nullptr, d_alm,
                                 st);
}

// ring-Fourier coefficients G (d_FG) -> pixels: out[pixel][nrp] = residual ? map - f : f
static int run_ring_synthesis(ShtPlan* p, const double* map, int64_t ldm, int residual, double* out, cudaStream_t st) {
    RingTabs rt{p->d_nphi.p, p->d_start.p, p->d_shift.p, p->d_twoff.p, p->d_tw.p};
    if (p->ntiles > 0) {
        dim3 gr(p->ntiles, (unsigned)ceil_div(p->nrp, 64));
        ring_synthesis_kernel<<<gr, kT, 0, st>>>(p->d_FG.p, rt, p->d_tile_ring.p, p->d_tile_j0.p, p->nrings,
                                                 p->lmax, p->nr, p->nrp, map, ldm, residual, out);
        SFB_CUDA_OK(cudaGetLastError());
        p->launches++;
    }
    if (p->n_ctiles > 0) {
        const int ni = pick_ni(p->nrp);
        dim3 gc(p->n_ctiles, (unsigned)ceil_div(p->nrp, 16 * ni));
#define SFB_LAUNCH_CS(NI_)                                                                                             \
do {                                                                                                               \
    constexpr int cs_bytes = cap_synthesis_smem_bytes<NI_>();                                                      \
    SFB_CUDA_OK(cudaFuncSetAttribute(cap_synthesis_kernel<NI_>, cudaFuncAttributeMaxDynamicSharedMemorySize,       \
                                     cs_bytes));                                                                   \
    cap_synthesis_kernel<NI_><<<gc, kT, cs_bytes, st>>>(p->d_FG.p, rt, p->d_ctile_ring.p, p->d_ctile_q0.p,         \
                                                        p->nrings, p->lmax, p->nr, p->nrp, map, ldm, residual,     \
                                                        out);                                             \
} while (0)
        if (ni == 4)
            SFB_LAUNCH_CS(4);
        else if (ni == 2)
            SFB_LAUNCH_CS(2);
        else
            SFB_LAUNCH_CS(1);
#undef SFB_LAUNCH_CS
        SFB_CUDA_OK(cudaGetLastError());
        p->launches++;
    }
    if (p->n_fft_rings > 0) {
        const int n = 4 * p->nside, sch = p->fft_sch;
        const int smem = n * sch * (int)sizeof(double2);
        SFB_CUDA_OK(
            cudaFuncSetAttribute(belt_synthesis_fft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        dim3 gf(p->n_fft_rings, (unsigned)ceil_div(p->nrp, 2 * sch));
        belt_synthesis_fft_kernel<<<gf, kT, smem, st>>>(p->d_FG.p, rt, p->d_fft_rings.p, p->nrings, p->lmax, p->log2n,
                                                        sch, p->nr, p->nrp, map, ldm, residual, out);
        SFB_CUDA_OK(cudaGetLastError());
        p->launches++;
    }
    return 0;
}

int sht_map2alm(ShtPlan* p, const double* d_win, int64_t ldw, int niter, double* d_alm, cudaStream_t st) {
    SFB_REQUIRE(p && d_win && d_alm, "sht_map2alm: null pointer");
    SFB_REQUIRE(ldw >= p->nr, "sht_map2alm: ld_win < nr");
    SFB_REQUIRE(niter >= 0, "sht_map2alm: niter < 0");
    cudaEvent_t e0, e1;
    SFB_CUDA_OK(cudaEventCreate(&e0));
    SFB_CUDA_OK(cudaEventCreate(&e1));
    SFB_CUDA_OK(cudaEventRecord(e0, st));
    p->launches = 0;
    const double* map = d_win;
    int64_t ldm = ldw;
    if (p->nside_in != p->nside) {
        udgrade_kernel<<<2048, 256, 0, st>>>(d_win, ldw, p->nside_in, p->d_map.p, p->nside, p->nr, p->nrp);
        SFB_CUDA_OK(cudaGetLastError());
        p->launches++;
        map = p->d_map.p;
        ldm = p->nrp;
    }
    RingTabs rt{p->d_nphi.p, p->d_start.p, p->d_shift.p, p->d_twoff.p, p->d_tw.p};
    const double w = 4.0 * 3.14159265358979323846 / (double)p->npix;
    const bool pixel_iter = getenv("SFB_SHT_PIXEL_ITER") != nullptr;  // cross-check: refine through pixel space
    auto legendre_synthesis = [&](bool ring_iter, int kend) -> int {
        double* f2 = ring_iter ? p->d_F2.p : nullptr;
        const int* mt = ring_iter ? p->d_mlim_tile.p : nullptr;
        const int nil = pick_ni(2 * p->nrp);
        dim3 gs(p->lmax + 1, (unsigned)ceil_div(kend, 64), (unsigned)ceil_div(2 * p->nrp, 16 * nil));
        if (nil == 4)
            legendre_synthesis_kernel<4><<<gs, kT, 0, st>>>(d_alm, p->d_lam.p, p->nrings, p->nhalf, kend, p->lmax, p->nrp,
                                                            p->d_FG.p, f2, p->d_nphi.p, mt);
        else if (nil == 2)
            legendre_synthesis_kernel<2><<<gs, kT, 0, st>>>(d_alm, p->d_lam.p, p->nrings, p->nhalf, kend, p->lmax, p->nrp,
                                                            p->d_FG.p, f2, p->d_nphi.p, mt);
        else
            legendre_synthesis_kernel<1><<<gs, kT, 0, st>>>(d_alm, p->d_lam.p, p->nrings, p->nhalf, kend, p->lmax, p->nrp,
                                                            p->d_FG.p, f2, p->d_nphi.p, mt);
        SFB_CUDA_OK(cudaGetLastError());
        p->launches += 1;
        return 0;
    };
    if (niter > 0 && !pixel_iter) {
        // alm <- alm + A f - Λ F'(Λᵀ alm): the refinement never leaves ring-Fourier space (ring_alias_kernel)
        const size_t nalm = p->lmsize * 2 * p->nrp;
        if (!p->d_F2.p) {
            SFB_TRY(p->d_F2.alloc((size_t)(p->lmax + 1) * p->nrings * 2 * p->nrp));
            SFB_CUDA_OK(cudaMemsetAsync(p->d_F2.p, 0, p->d_F2.n * sizeof(double), st));   // see the m-cutoff: finite from the start
        }
        SFB_TRY(p->d_a0.alloc(nalm));
        SFB_TRY(run_ring_analysis(p, map, ldm, st));
        SFB_TRY(run_legendre_analysis(p, p->d_FG.p, w, 0, nullptr, p->d_a0.p, st));
        SFB_CUDA_OK(cudaMemcpyAsync(d_alm, p->d_a0.p, nalm * sizeof(double), cudaMemcpyDeviceToDevice, st));
        const bool gram = p->kpolar < p->nhalf;
        if (gram) SFB_TRY(p->d_t.alloc(nalm));
        for (int it = 0; it < niter; ++it) {
            const double* add = p->d_a0.p;
            bool gforked = false;
            if (gram) {   // t = A f - K alm  (alias-free rings), from the alm of the previous pass
                // independent of the polar synthesis + alias pass below (both read alm, write different buffers): side stream
                gforked = p->kpolar > 0 && sht_fork(p, st);
                cudaStream_t gs = gforked ? p->side : st;
                const int nil = pick_ni(2 * p->nrp);
                dim3 gg((unsigned)ceil_div(p->lmax / 2 + 1, 64), (unsigned)ceil_div(2 * p->nrp, 16 * nil),
                        2 * (unsigned)(p->lmax + 1));
                if (nil == 4)
                    gram_apply_kernel<4><<<gg, kT, 0, gs>>>(p->d_gram.p, p->d_gram_off.p, d_alm, p->d_a0.p, p->lmax, p->nrp,
                                                            p->d_t.p);
                else if (nil == 2)
                    gram_apply_kernel<2><<<gg, kT, 0, gs>>>(p->d_gram.p, p->d_gram_off.p, d_alm, p->d_a0.p, p->lmax, p->nrp,
                                                            p->d_t.p);
                else
                    gram_apply_kernel<1><<<gg, kT, 0, gs>>>(p->d_gram.p, p->d_gram_off.p, d_alm, p->d_a0.p, p->lmax, p->nrp,
                                                            p->d_t.p);
                SFB_CUDA_OK(cudaGetLastError());
                p->launches++;
                add = p->d_t.p;
            }
            if (p->kpolar > 0) {
                SFB_TRY(legendre_synthesis(true, p->kpolar));
                const int alias_smem = (p->lmax + 1) * 2 * kAliasCW * (int)sizeof(double);
                if (p->n_alias_rings > 0 && alias_smem <= 200 * 1024 && !getenv("SFB_ALIAS_OLD")) {
                    SFB_CUDA_OK(cudaFuncSetAttribute(ring_alias_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                     alias_smem));
                    ring_alias_smem_kernel<<<dim3(p->n_alias_rings, (unsigned)ceil_div(p->nrp, kAliasCW)), 256, alias_smem,
                                             st>>>(p->d_FG.p, p->d_F2.p, p->d_alias_rings.p, p->d_nphi.p, p->d_shift.p,
                                                   p->nrings, p->lmax, p->nrp, p->d_mlim_ring.p);
                } else if (p->n_alias_rings > 0) {
                    ring_alias_kernel<<<dim3(p->nrings, p->lmax + 1), 64, 0, st>>>(p->d_FG.p, p->d_F2.p, p->d_nphi.p,
                                                                                  p->d_shift.p, p->nrings, p->lmax, p->nrp);
                }
                SFB_CUDA_OK(cudaGetLastError());
                p->launches++;
                if (gforked) SFB_TRY(sht_join(p, st));   // t is complete before the analysis adds it
                SFB_TRY(run_legendre_analysis(p, p->d_F2.p, -w, 1, add, d_alm, st, p->kpolar));
            }
        }
    } else {
    SFB_TRY(run_analysis(p, map, ldm, 0, d_alm, st));
    if (niter > 0) SFB_TRY(p->d_resid.alloc((size_t)p->npix * p->nrp));
    for (int it = 0; it < niter; ++it) {
        SFB_TRY(legendre_synthesis(false, p->nhalf));
        SFB_TRY(run_ring_synthesis(p, map, ldm, 1, p->d_resid.p, st));
        SFB_TRY(run_analysis(p, p->d_resid.p, p->nrp, 1, d_alm, st));
    }
    }
    SFB_CUDA_OK(cudaEventRecord(e1, st));
    if (p->pending) {
        cudaEventDestroy(p->tev0);
        cudaEventDestroy(p->tev1);
    }
    p->tev0 = e0;
    p->tev1 = e1;
    p->pending = true;
    return p->async_times ? 0 : sht_resolve_times(p);
}

// alm2map for all shells of the plan: planar alm -> d_out[pixel][nrp] (Healpix.alm2map! per shell, src/cat2anlm.jl:415)
int sht_alm2map(ShtPlan* p, const double* d_alm, double* d_out, cudaStream_t st) {
    SFB_REQUIRE(p && d_alm && d_out, "sht_alm2map: null pointer");
    const int nil = pick_ni(2 * p->nrp);
    dim3 gs(p->lmax + 1, (unsigned)ceil_div(p->nhalf, 64), (unsigned)ceil_div(2 * p->nrp, 16 * nil));
    if (nil == 4)
        legendre_synthesis_kernel<4><<<gs, kT, 0, st>>>(d_alm, p->d_lam.p, p->nrings, p->nhalf, p->nhalf, p->lmax, p->nrp,
                                                        p->d_FG.p, nullptr, p->d_nphi.p, nullptr);
    else if (nil == 2)
        legendre_synthesis_kernel<2><<<gs, kT, 0, st>>>(d_alm, p->d_lam.p, p->nrings, p->nhalf, p->nhalf, p->lmax, p->nrp,
                                                        p->d_FG.p, nullptr, p->d_nphi.p, nullptr);
    else
        legendre_synthesis_kernel<1><<<gs, kT, 0, st>>>(d_alm, p->d_lam.p, p->nrings, p->nhalf, p->nhalf, p->lmax, p->nrp,
                                                        p->d_FG.p, nullptr, p->d_nphi.p, nullptr);
    SFB_CUDA_OK(cudaGetLastError());
    p->launches = 1;
    return run_ring_synthesis(p, d_out, p->nrp, 0, d_out, st);
}

int sht_resolve_times(ShtPlan* p) {
    if (!p || !p->pending) return 0;
    SFB_CUDA_OK(cudaEventSynchronize(p->tev1));
    cudaEventElapsedTime(&p->t_total, p->tev0, p->tev1);
    cudaEventDestroy(p->tev0);
    cudaEventDestroy(p->tev1);
    p->tev0 = p->tev1 = nullptr;
    p->pending = false;
    return 0;
}

int alm_planar_to_complex(const double* d_alm, int lmax, int nr, int nrp, int layout, double* d_out, cudaStream_t st) {
    SFB_REQUIRE(d_alm && d_out, "alm_planar_to_complex: null pointer");
    alm_to_complex_kernel<<<dim3(lmax + 1, lmax + 1), 64, 0, st>>>(d_alm, lmax, nr, nrp, layout, d_out);
    SFB_CUDA_OK(cudaGetLastError());
    return 0;
}

int sht_alm_to_complex(const ShtPlan* p, const double* d_alm, int layout, double* d_out, cudaStream_t st) {
    SFB_REQUIRE(p, "sht_alm_to_complex: null pointer");
    return alm_planar_to_complex(d_alm, p->lmax, p->nr, p->nrp, layout, d_out, st);
}

}  // namespace sfb
