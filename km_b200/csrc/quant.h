// Stage 3 of km_find_batch: naming, clustering and FP64 quantification of one target's
// alternative paths by the CTA that built its graph.
//
//   diff_paths      MutationFinder.diff_path_without_overlap   (MutationFinder.py:321-373)
//   classify        MutationFinder.get_name                     (:429-488)
//   gram_columns, solve_wide, quant_pair   PathQuant.__init__/compute_coef/refine_coef/get_ratio (PathQuant.py:94-149)
//   emit_rows       quantify_paths (:613-648), _find_clusters (:651-723),
//                   quantify_clusters (:749-811)
//
// Every scan over a path (common prefix / suffix, minimum count, contribution sums) is done by
// the whole CTA: lanes test strided positions and combine with one shared-memory atomic, so the
// index pool in HBM is read with independent loads instead of a loop-carried chain.
//
// The least-squares step works on the normal equations: G = A^T A and h = A^T b are sums of
// small integers and float32-exact counts, accumulated in 64-bit integers, hence EXACT; the
// minimum-norm solution (what LAPACK gelsd returns for np.linalg.lstsq, PathQuant.py:116) comes
// from a Jacobi eigen-decomposition of G.  refine_coef is then iterated literally (fixed step
// 0.1, gradient / n_nodes, stop at max|grad| <= 0.01): the reference's answer IS the iterate it
// stops at.
#pragma once
#include "graph.h"

namespace km {

KM_HD int pv_at(const PathView& p, int i) {
    const int q = p.begin + i;
    if (p.bub_nk >= 0) return q <= p.bub_a ? q : (q <= p.bub_a + p.bub_nk ? p.idx[q - p.bub_a - 1] : p.bub_b + (q - p.bub_a - 1 - p.bub_nk));
    return p.c16 ? (int)p.c16[q] : (p.idx ? p.idx[q] : q);
}
// path p of the target (rank order) as a view: its node numbers come from the shared-memory cache when it holds them
// (graph_target leaves ce_len[p] = length and ce_b[p] = cache offset or -1)
KM_HD PathView path_view(const GraphScratch& S, const ResultView& R, int first_path, int p) {
    const int at = S.ce_b[p];
    PathView v = range_view(0, S.ce_len[p]);
    if (at == -2) return v;                              // the reference path of a simple bubble: position q holds node q
    if (at == -3) { v.idx = S.cand; v.bub_a = S.ce_a[0]; v.bub_nk = S.ce_a[1]; v.bub_b = S.ce_a[2]; return v; }
    v.c16 = (S.pcache_cap > 0 && at >= 0) ? S.pcache + at : nullptr;
    v.idx = R.pool + R.path_off[first_path + p];
    return v;
}
// Python indexing: a negative position counts from the end (the reference's third scan can run
// its alt cursor below zero, MutationFinder.py:362-369); beyond that Python raises -> no match
KM_HD int pv_at_wrap(const PathView& p, int i) {
    if (i < 0) i += p.len;
    return i < 0 ? -1 : pv_at(p, i);
}

struct Diff {
    int start, end_ref, end_var, end_ref_overlap;
};

// smallest p in [0, n) with f(p), else n.  All threads call; `slot` is CTA-shared.
template <class Ctx, class F>
KM_HD int first_true(const Ctx& ctx, int n, F f, int* slot) {
    if (ctx.tid() == 0) *slot = n;
    ctx.sync();
    for (int p = ctx.tid(); p < n; p += ctx.nt())
        if (f(p)) { atomic_mini32(slot, p); break; }
    ctx.sync();
    const int r = *slot;
    ctx.sync();
    return r;
}

template <class Ctx>
KM_HD Diff diff_paths(const Ctx& ctx, const PathView& ref, const PathView& alt, int k, int* slot) {
    const int nr = ref.len, na = alt.len, m = nr < na ? nr : na;
    if (!alt.idx && alt.bub_nk < 0 && !ref.idx && ref.bub_nk < 0 && alt.begin == ref.begin && na == nr) {
        // the reference against itself: the prefix scan runs to the end, the two suffix scans have no room
        Diff same;
        same.start = nr; same.end_ref = nr; same.end_var = nr; same.end_ref_overlap = nr;
        return same;
    }
    // common prefix (:321-331)
    const int i = first_true(ctx, m, [&](int p) { return pv_at(ref, p) != pv_at(alt, p); }, slot);
    // common suffix, keeping k positions clear of the prefix (:334-356): the loop runs while
    // both ends stay >= i + k, i.e. for at most m - i - k + 1 steps
    int room = m - i - k + 1;
    if (room < 0) room = 0;
    const int s1 = first_true(ctx, room, [&](int s) { return pv_at(ref, nr - 1 - s) != pv_at(alt, na - 1 - s); }, slot);
    const int jr = nr - s1, ja = na - s1;
    // the same scan allowed to run back to the prefix on the reference side (:358-369)
    const int s2 = first_true(ctx, jr - i, [&](int s) { return pv_at(ref, jr - 1 - s) != pv_at_wrap(alt, ja - 1 - s); }, slot);
    Diff d;
    d.start = i; d.end_ref = jr; d.end_var = ja; d.end_ref_overlap = jr - s2;
    return d;
}

enum { KM_T_REFERENCE = 0, KM_T_SUBSTITUTION = 1, KM_T_ITD = 2, KM_T_INDEL = 3, KM_T_INSERTION = 4, KM_T_DELETION = 5 };

// get_name: type + trimmed deleted / inserted runs.  kmers[i] = last base of canonical node i (all the naming
// needs of a k-mer: the scan below is serial, one lane, and used to wait for an L2 round trip per step).
// (`cut` = length of the suffix the deleted and the inserted string share, when the caller has it already: shared_suffix)
KM_HD int classify(const uint8_t* kmers, const PathView& ref, const PathView& alt, const Diff& d,
                   int* del_len, int* ins_len, int cut_known = -1) {
    int gone = d.end_ref - d.start, fresh = d.end_var - d.start;
    int cut = 0;
    if (cut_known >= 0) cut = cut_known;
    else if (gone > 0) {     // strip the suffix both strings share (:446-456)
        while (cut < gone && cut < fresh &&
               kmers[pv_at(ref, d.end_ref - 1 - cut)] == kmers[pv_at(alt, d.end_var - 1 - cut)])
            ++cut;
    }
    gone -= cut; fresh -= cut;
    *del_len = gone; *ins_len = fresh;
    if (d.end_ref == d.end_var) return d.start == d.end_ref ? KM_T_REFERENCE : KM_T_SUBSTITUTION;
    if (d.start == d.end_ref_overlap) return KM_T_ITD;
    if (d.end_ref < d.end_var) return gone == 0 ? KM_T_INSERTION : KM_T_INDEL;
    return fresh == 0 ? KM_T_DELETION : KM_T_INDEL;
}

// The scan of classify by a whole warp (all lanes call): lanes compare strided positions, the first mismatch wins.
template <class WCtx>
KM_HD int shared_suffix(const WCtx& wctx, const uint8_t* kmers, const PathView& ref, const PathView& alt, const Diff& d, int* slot) {
    const int gone = d.end_ref - d.start, fresh = d.end_var - d.start;
    const int m = gone > 0 ? (gone < fresh ? gone : fresh) : 0;
    return first_true(wctx, m, [&](int s) { return kmers[pv_at(ref, d.end_ref - 1 - s)] != kmers[pv_at(alt, d.end_var - 1 - s)]; }, slot);
}

#if KM_DEVICE_BUILD
#define KM_COLD __device__ __noinline__
#else
#define KM_COLD static
#endif

// cyclic Jacobi on the symmetric m x m matrix A (destroyed); V receives the eigenvectors
// (columns), the diagonal of A the eigenvalues.
KM_COLD void jacobi_eigen(double* A, double* V, int m) {
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < m; ++j) V[i * m + j] = i == j ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0.0, diag = 0.0;
        for (int i = 0; i < m; ++i) {
            diag += A[i * m + i] * A[i * m + i];
            for (int j = i + 1; j < m; ++j) off += A[i * m + j] * A[i * m + j];
        }
        if (off <= 1e-34 * diag || off == 0.0) break;
        for (int p = 0; p < m - 1; ++p)
            for (int q = p + 1; q < m; ++q) {
                const double apq = A[p * m + q];
                if (apq == 0.0) continue;
                const double theta = (A[q * m + q] - A[p * m + p]) / (2.0 * apq);
                const double tt = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(tt * tt + 1.0), s = tt * c;
                for (int r = 0; r < m; ++r) {
                    const double arp = A[r * m + p], arq = A[r * m + q];
                    A[r * m + p] = c * arp - s * arq;
                    A[r * m + q] = s * arp + c * arq;
                }
                for (int r = 0; r < m; ++r) {
                    const double apr = A[p * m + r], aqr = A[q * m + r];
                    A[p * m + r] = c * apr - s * aqr;
                    A[q * m + r] = s * apr + c * aqr;
                }
                for (int r = 0; r < m; ++r) {
                    const double vrp = V[r * m + p], vrq = V[r * m + q];
                    V[r * m + p] = c * vrp - s * vrq;
                    V[r * m + q] = s * vrp + c * vrq;
                }
            }
    }
}

// a*d - b*c with one rounding of the exact result's neighbourhood (Kahan); inputs are integers
// below 2^53 held in doubles, so the only error is the final ~1.5 ulp
KM_HD double det2(double a, double b, double c, double d) {
    const double w = b * c;
    const double e = fma(-b, c, w);
    const double f = fma(a, d, -w);
    return f + e;
}

// ---- refine_coef without its thousands of idle iterations ----------------------------------------------
// The reference iterates coef += 0.1 * grad with grad = 2 (h - G coef) / n until max|grad| <= 0.01
// (PathQuant.py:120-142).  When lstsq leaves a negative coefficient (a cluster whose second variant explains
// nothing) the clamped problem converges along the slowest eigen-direction of G and the loop runs hundreds to
// thousands of times -- one such target then takes longer than the rest of the batch.  Between two clamp
// events the iteration is LINEAR in the free coefficients: x(t+1) = x(t) + a (h_F - G_FF x(t)), a = 0.2/n,
// so in the eigenbasis of G_FF every component is y_e(t) = y_e(0) r_e^t + g_e (1 - r_e^t) / l_e with
// r_e = 1 - a l_e in (0, 1], and the gradient component is d_e(t) = (g_e - l_e y_e(0)) r_e^t.  refine_jump
// evaluates that closed form to (1) predict the step at which the stop test will fire and (2) advance the
// state to 8 steps before it -- but only over a stretch [0, J] on which it can PROVE that the literal loop
// does nothing else: every y_e(t) and d_e(t) is monotone in t, so on a segment [t0, t1] each term V_ie y_e(t)
// lies between its two end values and
//     min over the segment of x_i(t)   >=  sum_e min(V_ie y_e(t0), V_ie y_e(t1))          (no free coefficient is clamped)
//     max over the segment of grad_a(t) <=  c_a - sum_e min(w_ae y_e(t0), w_ae y_e(t1))    (no clamped one is released)
//     min over the segment of |grad_i|  >=  the same bound on sum_e V_ie d_e(t)            (the stop test does not fire)
// hold for every real t in the segment, hence for every step.  [0, J] is cut adaptively (a segment whose bounds
// are inconclusive is halved, down to single steps); if the proof fails, J is halved; below 32 steps the literal
// loop simply goes on.  The literal loop then runs the last steps and decides the stop itself, so the result
// differs from the fully literal run only by the rounding of the closed form (~1e-12 relative), and the first 32
// iterations (every transient, every case the reference's own tests hold) are literal.
//
// WARP-COLLECTIVE: all `nl` lanes of the calling warp pass the same arguments (each its own copy, or memory nobody
// writes during the call) and get the same result.  The closed form costs two exponentials per mode, and on one lane
// the search for the stop (doubling + bisection) and the proof were ~20 evaluations one after the other -- the longest
// serial stretch of a whole batch.  Here 32 candidate steps are evaluated at once, one per lane (a 33-ary search), and
// the proof runs on 32 sub-stretches side by side.  (The CPU emulation's single lane takes the candidates in turn.)
#define KM_REFINE_MAXF 4
// (not inlined: it runs for a handful of targets per batch and must not cost the graph kernel its registers)
KM_COLD int refine_jump(int lane, int nl, const double* G, const double* h, int m, int n_nodes, const double* coef, double* coef_new) {
    int F[KM_REFINE_MAXF], mf = 0;
    for (int a = 0; a < m; ++a) coef_new[a] = coef[a];
    for (int a = 0; a < m; ++a)
        if (coef[a] > 0.0) { if (mf == KM_REFINE_MAXF) return 0; F[mf++] = a; }
    if (mf == 0 || m > 2 * KM_REFINE_MAXF) return 0;
    double A[KM_REFINE_MAXF * KM_REFINE_MAXF], V[KM_REFINE_MAXF * KM_REFINE_MAXF];
    double lam[KM_REFINE_MAXF], lg[KM_REFINE_MAXF], y0[KM_REFINE_MAXF], g[KM_REFINE_MAXF], d0[KM_REFINE_MAXF];
    for (int i = 0; i < mf; ++i)
        for (int j = 0; j < mf; ++j) A[i * mf + j] = G[F[i] * m + F[j]];
    jacobi_eigen(A, V, mf);
    const double alpha = 0.2 / (double)n_nodes;
    for (int e = 0; e < mf; ++e) {
        lam[e] = A[e * mf + e] > 0.0 ? A[e * mf + e] : 0.0;
        if (!(alpha * lam[e] < 1.0)) return 0;                 // an oscillating mode: leave it to the literal loop
        lg[e] = log1p(-alpha * lam[e]);                        // log r_e <= 0
        y0[e] = 0.0; g[e] = 0.0;
        for (int i = 0; i < mf; ++i) { y0[e] += V[i * mf + e] * coef[F[i]]; g[e] += V[i * mf + e] * h[F[i]]; }
        d0[e] = g[e] - lam[e] * y0[e];
    }
    // clamped coefficients: grad_a(t) * n / 2 = h_a - sum_e w_ae y_e(t)
    int Cl[2 * KM_REFINE_MAXF], mc = 0;
    double wc[2 * KM_REFINE_MAXF * KM_REFINE_MAXF];
    for (int a = 0; a < m; ++a) {
        bool is_free = false;
        for (int i = 0; i < mf; ++i) is_free |= F[i] == a;
        if (is_free) continue;
        for (int e = 0; e < mf; ++e) {
            double w = 0.0;
            for (int i = 0; i < mf; ++i) w += G[a * m + F[i]] * V[i * mf + e];
            wc[mc * KM_REFINE_MAXF + e] = w;
        }
        Cl[mc++] = a;
    }
    // y_e and d_e after t steps
#if defined(KM_HOST_EMU) && defined(KM_RJ_DEBUG)
    int n_eval = 0;
#endif
    auto eval = [&](double t, double* y, double* d) {
#if defined(KM_HOST_EMU) && defined(KM_RJ_DEBUG)
        ++n_eval;
#endif
        for (int e = 0; e < mf; ++e) {
            const double tl = t * lg[e];
            const double rt = exp(tl);
            const double grow = lam[e] > 0.0 ? -expm1(tl) / lam[e] : alpha * t;      // (1 - r^t) / l
            y[e] = y0[e] * rt + g[e] * grow;
            d[e] = d0[e] * rt;
        }
    };
    const double gscale = 2.0 / (double)n_nodes;
    // does the stop test fire at step t?  (max |grad| over the free coefficients <= 0.01)
    auto quiet_at = [&](double t) {
        double y[KM_REFINE_MAXF], d[KM_REFINE_MAXF];
        eval(t, y, d);
        double w = 0.0;
        for (int i = 0; i < mf; ++i) {
            double gi = 0.0;
            for (int e = 0; e < mf; ++e) gi += V[i * mf + e] * d[e];
            const double ag = fabs(gi) * gscale;
            w = ag > w ? ag : w;
        }
        return w <= 0.01;
    };
    // the three interval bounds on [tA, tB] from the end states; true = nothing but linear steps in between
    auto segment_clean = [&](const double* yA, const double* dA, const double* yB, const double* dB) {
        for (int i = 0; i < mf; ++i) {
            double lo = 0.0, mag = 0.0;
            for (int e = 0; e < mf; ++e) {
                const double a = V[i * mf + e] * yA[e], b = V[i * mf + e] * yB[e];
                lo += a < b ? a : b;
                mag += fabs(a) > fabs(b) ? fabs(a) : fabs(b);
            }
            if (!(lo > 1e-9 * mag)) return false;
        }
        for (int c = 0; c < mc; ++c) {
            double lo = 0.0, mag = fabs(h[Cl[c]]);
            for (int e = 0; e < mf; ++e) {
                const double a = wc[c * KM_REFINE_MAXF + e] * yA[e], b = wc[c * KM_REFINE_MAXF + e] * yB[e];
                lo += a < b ? a : b;
                mag += fabs(a) > fabs(b) ? fabs(a) : fabs(b);
            }
            if (!(h[Cl[c]] - lo < -1e-9 * mag)) return false;
        }
        bool loud = false;                       // some free coefficient's |grad| stays above the stop threshold
        for (int i = 0; i < mf && !loud; ++i) {
            double lo = 0.0, hi = 0.0;
            for (int e = 0; e < mf; ++e) {
                const double a = V[i * mf + e] * dA[e], b = V[i * mf + e] * dB[e];
                lo += a < b ? a : b;
                hi += a < b ? b : a;
            }
            loud = lo * gscale > 0.01 * (1.0 + 1e-9) || hi * gscale < -0.01 * (1.0 + 1e-9);
        }
        return loud;
    };
    // [t0, t1] proven clean by adaptive bisection (explicit stack; the terms are exponentials that have shed their fast
    // modes during the 32 literal steps before the call, so a handful of leaves is the rule)
    auto stretch_clean = [&](double t0, double t1) {
        double st_lo[24], st_hi[24];
        int sp = 0, work = 0;
        st_lo[0] = t0; st_hi[0] = t1; sp = 1;
        while (sp) {
            --sp;
            const double a = st_lo[sp], b = st_hi[sp];
            double yA[KM_REFINE_MAXF], dA[KM_REFINE_MAXF], yB[KM_REFINE_MAXF], dB[KM_REFINE_MAXF];
            eval(a, yA, dA); eval(b, yB, dB);
            if (segment_clean(yA, dA, yB, dB)) continue;
            if (b - a <= 1.0 || sp + 2 > 24 || ++work > 64) return false;
            const double mid = floor(0.5 * (a + b));
            st_lo[sp] = mid; st_hi[sp] = b; ++sp;
            st_lo[sp] = a; st_hi[sp] = mid; ++sp;
        }
        return true;
    };
    // (1) the first power-of-two multiple of 32 steps at which the stop test fires: 32 * 2^v, v = 0..18, a lane each.
    // (Only a PREDICTION: whatever it says, the stretch actually jumped is proven clean below and the literal loop
    // decides the stop.)
    uint32_t quiet = 0;
    for (int v = lane; v < 19; v += nl)
        if (quiet_at(ldexp(32.0, v))) { quiet |= 1u << v; if (nl == 1) break; }
    quiet = warp_or32(quiet);
    if (!quiet) return 0;
    const int v0 = ffs32(quiet) - 1;
    double hi = ldexp(32.0, v0), lo = v0 ? 0.5 * hi : 0.0;
    // (2) the first step in (lo, hi] at which it fires, 32 candidates per round
    while (hi - lo > 1.0) {
        const double width = hi - lo;
        const double stride = width <= 33.0 ? 1.0 : ceil(width / 33.0);
        uint32_t fired = 0, asked = 0;
        for (int v = lane; v < 32; v += nl) {
            const double t = lo + stride * (double)(v + 1);
            if (t < hi) { asked |= 1u << v; if (quiet_at(t)) fired |= 1u << v; }
        }
        fired = warp_or32(fired); asked = warp_or32(asked);
        if (fired) { const int v = ffs32(fired) - 1; hi = lo + stride * (double)(v + 1); lo = hi - stride; }
        else lo = lo + stride * (double)popc32(asked);
    }
    // (3) the longest stretch [0, J], J = hi - 8, hi/2 - 4, ... that is provably clean: 32 sub-stretches side by side
    double J = hi - 8.0;
    for (; J >= 32.0; J = floor(0.5 * J)) {
        bool clean = true;
        for (int v = lane; v < 32 && clean; v += nl) {
            const double a = floor(J * (double)v / 32.0), b = v == 31 ? J : floor(J * (double)(v + 1) / 32.0);
            if (b > a) clean = stretch_clean(a, b);
        }
        if (!warp_or32(clean ? 0u : 1u)) break;
    }
#if defined(KM_HOST_EMU) && defined(KM_RJ_DEBUG)
    fprintf(stderr, "refine_jump: mf=%d predicted stop %.0f, proven stretch %.0f, %d closed-form evaluations\n", mf, hi, J, n_eval);
#endif
    if (J < 32.0) return 0;
    double y[KM_REFINE_MAXF], d[KM_REFINE_MAXF];
    eval(J, y, d);
    for (int i = 0; i < mf; ++i) {
        double x = 0.0;
        for (int e = 0; e < mf; ++e) x += V[i * mf + e] * y[e];
        coef_new[F[i]] = x;
    }
    return (int)J;
}

// refine_coef (PathQuant.py:120-136) + get_ratio (:144-149) on the normal equations: fixed step 0.1,
// gradient / n_nodes, stop at max|grad| <= 0.01 -- literal steps, with refine_jump across the long
// linear stretches.  Returns the number of iterations the literal loop would have run.
// M columns, known at compile time: the whole state lives in registers (in shared memory one step of a 3-column
// problem was a 1,650-cycle chain of dependent loads, stores and divisions; in registers the rows are independent).
// WARP-COLLECTIVE like refine_jump: every lane computes the same thing, which costs nothing and lets the jump use them.
// G_in / h_in must stay valid and unchanged during the call.
// (G is symmetric: its M (M + 1) / 2 distinct entries are kept, which is what lets three columns run without spills in a
// 64-register kernel)
KM_HD constexpr int sym_index(int a, int b) { return a >= b ? a * (a + 1) / 2 + b : b * (b + 1) / 2 + a; }
template <int M>
KM_HD int refine_and_ratio(int lane, int nl, const double* G_in, const double* h_in, int n_nodes, double* coef_io, double* rvaf,
                            bool allow_jump) {
    double G[M * (M + 1) / 2], h[M], c[M], grad[M];
#pragma unroll
    for (int a = 0; a < M; ++a)
#pragma unroll
        for (int b = 0; b <= a; ++b) G[sym_index(a, b)] = G_in[a * M + b];
#pragma unroll
    for (int a = 0; a < M; ++a) { h[a] = h_in[a]; c[a] = coef_io[a] < 0.0 ? 0.0 : coef_io[a]; }
    double worst = INFINITY;
    int iters = 0, since = 0;
    while (worst > 0.01) {
#pragma unroll
        for (int a = 0; a < M; ++a) {
            double fit = 0.0;
#pragma unroll
            for (int b = 0; b < M; ++b) fit += G[sym_index(a, b)] * c[b];
            grad[a] = 2.0 * (h[a] - fit) / (double)n_nodes;
        }
        worst = 0.0;
#pragma unroll
        for (int a = 0; a < M; ++a) {
            c[a] += 0.1 * grad[a];
            if (c[a] < 0.0) { grad[a] = 0.0; c[a] = 0.0; }
            const double ag = fabs(grad[a]);
            worst = ag > worst ? ag : worst;     // NaN never enters: counts are finite
        }
        if (++iters > 10000000) { iters = -1; break; }
        if (allow_jump && ++since >= 32 && worst > 0.01) {
#if KM_DEVICE_BUILD && defined(KM_PHASE_TIMERS)
            const long long j0 = clock64();
#endif
            double cur[M], nxt[M];
#pragma unroll
            for (int a = 0; a < M; ++a) cur[a] = c[a];
            const int J = refine_jump(lane, nl, G_in, h_in, M, n_nodes, cur, nxt);
            if (J) {
#pragma unroll
                for (int a = 0; a < M; ++a) c[a] = nxt[a];
            }
            iters += J;
#if KM_DEVICE_BUILD && defined(KM_PHASE_TIMERS)
            if (lane == 0) atomicAdd(&km_phase_cycles[56], (unsigned long long)(clock64() - j0));
#endif
            since = 0;
        }
    }
    double cmax = c[0], csum = 0.0;
#pragma unroll
    for (int a = 0; a < M; ++a) { cmax = c[a] > cmax ? c[a] : cmax; csum += c[a]; }
#pragma unroll
    for (int a = 0; a < M; ++a) { coef_io[a] = c[a]; rvaf[a] = cmax == 0.0 ? c[a] : c[a] / csum; }
    return iters;
}

// (a function of its own for three and four columns: inside solve_wide the loop's state competed with everything else that
// is live there and spilled -- 700 bytes in the 64-register kernel, a local-memory round trip on every operand)
template <int M>
KM_COLD int refine_and_ratio_cold(int lane, int nl, const double* G_in, const double* h_in, int n_nodes, double* coef_io, double* rvaf,
                                  bool allow_jump) {
    return refine_and_ratio<M>(lane, nl, G_in, h_in, n_nodes, coef_io, rvaf, allow_jump);
}

// The same for any number of columns, state in memory (clusters of four variants and more): lane 0 iterates, the warp
// meets for the jumps.  coef / rvaf / grad: m doubles each, written by lane 0 only.
KM_HD int refine_and_ratio_any(int lane, int nl, const double* G, const double* h, int m, int n_nodes, double* coef, double* rvaf,
                                double* grad, bool allow_jump) {
    if (lane == 0) for (int a = 0; a < m; ++a) if (coef[a] < 0.0) coef[a] = 0.0;
    int iters = 0;
    bool more = true;
    while (more) {
        // literal steps until the stop, or until 32 of them have passed without it
        int go = 0;                                           // 0: stopped, 1: try a jump, 2: watchdog
        if (lane == 0) {
            double worst = INFINITY;
            int since = 0;
            while (worst > 0.01) {
                for (int a = 0; a < m; ++a) {
                    double fit = 0.0;
                    for (int b = 0; b < m; ++b) fit += G[a * m + b] * coef[b];
                    grad[a] = 2.0 * (h[a] - fit) / (double)n_nodes;
                }
                worst = 0.0;
                for (int a = 0; a < m; ++a) {
                    coef[a] += 0.1 * grad[a];
                    if (coef[a] < 0.0) { grad[a] = 0.0; coef[a] = 0.0; }
                    const double ag = fabs(grad[a]);
                    worst = ag > worst ? ag : worst;
                }
                if (++iters > 10000000) { go = 2; break; }
                if (allow_jump && m <= 2 * KM_REFINE_MAXF && ++since >= 32 && worst > 0.01) { go = 1; break; }
            }
        }
#if KM_DEVICE_BUILD
        __syncwarp();
#endif
        go = warp_shfl32(go, 0);
        iters = warp_shfl32(iters, 0);
        if (go == 2) { iters = -1; break; }
        more = go == 1;
        if (more) {
            double cur[2 * KM_REFINE_MAXF], nxt[2 * KM_REFINE_MAXF];
            for (int a = 0; a < m; ++a) cur[a] = coef[a];
            const int J = refine_jump(lane, nl, G, h, m, n_nodes, cur, nxt);
#if KM_DEVICE_BUILD
            __syncwarp();
#endif
            if (J && lane == 0) for (int a = 0; a < m; ++a) coef[a] = nxt[a];
            iters += J;
        }
    }
    if (lane == 0) {
        double cmax = coef[0], csum = 0.0;
        for (int a = 0; a < m; ++a) { cmax = coef[a] > cmax ? coef[a] : cmax; csum += coef[a]; }
        for (int a = 0; a < m; ++a) rvaf[a] = cmax == 0.0 ? coef[a] : coef[a] / csum;
    }
#if KM_DEVICE_BUILD
    __syncwarp();
#endif
    return iters;
}

// a + b exactly as a double-double (hi, lo)
KM_HD void two_sum(double a, double b, double* hi, double* lo) {
    const double s = a + b, bb = s - a;
    *hi = s; *lo = (a - (s - bb)) + (b - bb);
}
// sum_j c_j * x_j of three products to (almost) the last bit: products and partial sums carry their rounding errors along
KM_HD double dot3_compensated(const double* c, const double* x) {
    double hi = 0.0, lo = 0.0;
    for (int j = 0; j < 3; ++j) {
        const double p = c[j] * x[j];
        const double e = fma(c[j], x[j], -p);
        double s, r;
        two_sum(hi, p, &s, &r);
        hi = s; lo += r + e;
    }
    return hi + lo;
}

// The minimum-norm solution of the 3 x 3 normal equations G x = h from their exact integer sums (acc: lower triangle of G at
// [a * 3 + b], b <= a; h at [9..11]).  G = A^T A is positive semi-definite and h = A^T b lies in its range.
//   rank 3: x = adj(G) h / det(G), cofactors and determinant exact integers
//   rank 2: adj(G) = alpha n n^T with n spanning the null space (any non-zero column of the cofactor matrix); a particular
//           solution from the largest principal 2 x 2 minor (third unknown 0), minus its component along n
//   rank 1: G = lambda u u^T, lambda = trace; x = v (v.h) / (|v|^2 lambda) for any non-zero column v of G
// false: entries too large for exact 64-bit cofactors (the caller falls back to the eigen-decomposition).
KM_HD bool solve3_exact(const unsigned long long* acc, double* x) {
    long long a[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            const unsigned long long v = acc[(i >= j ? i : j) * 3 + (i >= j ? j : i)];
            if (v >= (1ull << 20)) return false;
            a[i * 3 + j] = (long long)v;
        }
    const double hh[3] = {(double)acc[9], (double)acc[10], (double)acc[11]};
    long long C[9];                       // cofactors (the matrix is symmetric, so is its adjugate)
    C[0] = a[4] * a[8] - a[5] * a[7]; C[1] = a[5] * a[6] - a[3] * a[8]; C[2] = a[3] * a[7] - a[4] * a[6];
    C[3] = C[1];                      C[4] = a[0] * a[8] - a[2] * a[6]; C[5] = a[1] * a[6] - a[0] * a[7];
    C[6] = C[2];                      C[7] = C[5];                      C[8] = a[0] * a[4] - a[1] * a[3];
    const long long det = a[0] * C[0] + a[1] * C[1] + a[2] * C[2];
    if (det != 0) {
        for (int i = 0; i < 3; ++i) {
            const double ci[3] = {(double)C[i * 3], (double)C[i * 3 + 1], (double)C[i * 3 + 2]};
            x[i] = dot3_compensated(ci, hh) / (double)det;
        }
        return true;
    }
    // the largest principal 2 x 2 minor (C[kk] = the minor that leaves row and column k out; >= 0 for a PSD matrix)
    int k = 0;
    if (C[4] > C[k * 4]) k = 1;
    if (C[8] > C[k * 4]) k = 2;
    if (C[k * 4] != 0) {
        const int i = k == 0 ? 1 : 0, j = k == 2 ? 1 : 2;
        const double gii = (double)a[i * 3 + i], gij = (double)a[i * 3 + j], gjj = (double)a[j * 3 + j], minor = (double)C[k * 4];
        double xp[3];
        xp[i] = det2(hh[i], hh[j], gij, gjj) / minor;            // Cramer on exact integers
        xp[j] = det2(gii, gij, hh[i], hh[j]) / minor;
        xp[k] = 0.0;
        const double n[3] = {(double)C[k * 3], (double)C[k * 3 + 1], (double)C[k * 3 + 2]};      // column k of the adjugate: null vector
        const double nn = n[0] * n[0] + n[1] * n[1] + n[2] * n[2];
        const double along = dot3_compensated(n, xp) / nn;
        for (int q = 0; q < 3; ++q) x[q] = xp[q] - n[q] * along;
        return true;
    }
    const long long trace = a[0] + a[4] + a[8];
    if (trace == 0) { x[0] = x[1] = x[2] = 0.0; return true; }
    int col = 0;
    if (a[4] > a[col * 4]) col = 1;
    if (a[8] > a[col * 4]) col = 2;
    const double v[3] = {(double)a[col], (double)a[3 + col], (double)a[6 + col]};
    const double sc = dot3_compensated(v, hh) / ((v[0] * v[0] + v[1] * v[1] + v[2] * v[2]) * (double)trace);
    for (int q = 0; q < 3; ++q) x[q] = v[q] * sc;
    return true;
}

// The serial part of a wide cluster's quantification, by warp 0 of the CTA (all its lanes call): G and h from the exact
// integer sums, the minimum-norm solution, refine_coef, get_ratio.  coef / rvaf (m doubles each, shared memory) are
// written by lane 0; returns the iteration count on every lane.  Not inlined: it runs for a handful of targets.
KM_COLD int solve_wide(int lane, int nl, const GraphScratch& S, int m, int n_nodes, double* coef, double* rvaf, bool allow_jump) {
    const unsigned long long* acc = S.acc;
    double* G = S.G;
    double* h = S.vec;   // [m]
    if (lane == 0) {
        for (int a = 0; a < m; ++a) {
            h[a] = (double)acc[m * m + a];
            for (int b = 0; b <= a; ++b) G[a * m + b] = G[b * m + a] = (double)acc[a * m + b];
        }
    }
#if KM_DEVICE_BUILD
    __syncwarp();
#endif
    PhaseTimer st;
    bool solved = false;
    if (m == 3) {
        // Three columns (a cluster of two variants, the common wide case): G is an exact integer matrix, so its rank is
        // decided exactly and the minimum-norm solution (what lstsq returns, PathQuant.py:116) has a closed form at every
        // rank -- no eigen-decomposition (33,000 cycles of one lane's square roots and divisions).  Two tandem copies of a
        // duplication next to one make the columns linearly dependent (2 * once - reference = twice): rank 2 is the rule.
        double c3[3];
        solved = solve3_exact(acc, c3);
        if (solved && lane == 0) { coef[0] = c3[0]; coef[1] = c3[1]; coef[2] = c3[2]; }
    }
    if (!solved) {
        if (lane == 0) {
            double* A = S.V;                 // working copy for the eigen solver
            double* V = S.V + m * m;         // needs 2*m*m doubles: S.V is sized for that
            for (int i = 0; i < m * m; ++i) A[i] = G[i];
            jacobi_eigen(A, V, m);
            double lmax = 0.0;
            for (int i = 0; i < m; ++i) lmax = A[i * m + i] > lmax ? A[i * m + i] : lmax;
            // singular values below eps*max(M,N)*s_max are dropped by lstsq(rcond=None); in
            // eigenvalue terms that is far below FP64 resolution of G, so the cut is placed where
            // an exactly rank-deficient integer G leaves its rounding noise.
            const double cut = lmax * 1e-11;
            for (int a = 0; a < m; ++a) coef[a] = 0.0;
            for (int e = 0; e < m; ++e) {
                const double lam = A[e * m + e];
                if (!(lam > cut)) continue;
                double proj = 0.0;
                for (int a = 0; a < m; ++a) proj += V[a * m + e] * h[a];
                proj /= lam;
                for (int a = 0; a < m; ++a) coef[a] += V[a * m + e] * proj;
            }
        }
    }
#if KM_DEVICE_BUILD
    __syncwarp();
#endif
    st.mark_warp0(54, lane);
    int iters;
    if (m == 3 || m == 4) {
        // state in registers, every lane the same
        double c[4], r[4];
        for (int a = 0; a < m; ++a) c[a] = coef[a];
#if KM_DEVICE_BUILD
        __syncwarp();
#endif
        iters = m == 3 ? refine_and_ratio_cold<3>(lane, nl, G, h, n_nodes, c, r, allow_jump)
                       : refine_and_ratio_cold<4>(lane, nl, G, h, n_nodes, c, r, allow_jump);
        if (lane == 0) for (int a = 0; a < m; ++a) { coef[a] = c[a]; rvaf[a] = r[a]; }
    } else {
        iters = refine_and_ratio_any(lane, nl, G, h, m, n_nodes, coef, rvaf, S.vec + 2 * S.max_cols, allow_jump);
    }
    st.mark_warp0(55, lane);
    return iters;
}

// The exact integer sums of m columns' normal equations into S.acc (solve_wide turns them into coefficients).  All threads
// of the CTA must call this; S.occ must be all zero on entry and is all zero again on return.
template <class Ctx>
KM_HD void gram_columns(const Ctx& ctx, const GraphScratch& S, const uint32_t* counts, const PathView* cols, int m) {
    const int tid = ctx.tid(), nt = ctx.nt();
    PhaseTimer st;                     // (measurement builds: 53 Gram sums, 54 eigen solve, 55 refine)
    unsigned long long* acc = S.acc;   // [m*m + m] exact integer accumulators
    for (int i = tid; i < m * m + m; i += nt) acc[i] = 0ull;
    ctx.sync();
    // contrib[i, c] = occurrences of node i in column c (PathQuant.py:101-104);
    // G[a][b] = sum_i contrib[i,a]*contrib[i,b]; h[a] = sum_i contrib[i,a]*float32(count_i).
    // A column that is a slice of the reference (idx == nullptr) touches each node of its range
    // once, so its occurrence vector is a range test; only real paths are scattered into S.occ.
    // Every lane sums privately, lanes combine by shuffle, and one lane per warp adds to acc.
    for (int b = 0; b < m; ++b) {
        const PathView cb = cols[b];
        unsigned long long hb = 0ull;
        for (int p = tid; p < cb.len; p += nt) {
            const int node = pv_at(cb, p);
            if (cb.idx) atomic_addi32(&S.occ[node], 1);
            hb += (unsigned long long)(float)counts[node];      // counts -> float32 (PathQuant.py:99), an integer
        }
        hb = warp_sum64(hb);
        if (warp_leader() && hb) atomic_add64(&acc[m * m + b], hb);
        ctx.sync();
        for (int a = b; a < m; ++a) {
            const PathView ca = cols[a];
            unsigned long long part = 0ull;
            for (int p = tid; p < ca.len; p += nt) {
                const int node = pv_at(ca, p);
                part += cb.idx ? (unsigned long long)S.occ[node] : (unsigned long long)(node >= cb.begin && node < cb.begin + cb.len);
            }
            part = warp_sum64(part);
            if (warp_leader() && part) atomic_add64(&acc[a * m + b], part);
        }
        ctx.sync();
        if (cb.idx) {
            for (int p = tid; p < cb.len; p += nt) S.occ[pv_at(cb, p)] = 0;
            ctx.sync();
        }
    }
    st.mark(53);
}

// min(counts over the path) by the whole group; `slot` is shared by the group.  An empty path gives 2^32-1.
template <class Ctx>
KM_HD int64_t min_count(const Ctx& ctx, const uint32_t* counts, const PathView& p, int* slot) {
    uint32_t* us = reinterpret_cast<uint32_t*>(slot);
    if (ctx.tid() == 0) *us = 0xFFFFFFFFu;
    ctx.sync();
    uint32_t mine = 0xFFFFFFFFu;
    for (int i = ctx.tid(); i < p.len; i += ctx.nt()) { const uint32_t c = counts[pv_at(p, i)]; mine = c < mine ? c : mine; }
    if (mine != 0xFFFFFFFFu) atomic_min32(us, mine);
    ctx.sync();
    const int64_t r = (int64_t)*us;
    ctx.sync();
    return r;
}

// Two-column quantification by ONE WARP: `path` is a real (possibly clipped) path, `range` a slice
// of the reference.  This is every vs_ref row ([alt, ref], MutationFinder.py:618-631) and every
// single-variant cluster ([ref_clip, alt_clip], :784-790).  The path's occurrence counts go into
// byte lane `lane8` of S.occ (a node occurs at most once on each of the two tree chains a path is
// stitched from), so the warps of a CTA work on different rows at the same time without CTA
// barriers; sums are combined by shuffles.  `path_first` says which column comes first in the
// reference's column order.  Returns refine iterations in *iters; coef/rvaf in reference order.
struct Quant2 {
    double coef[2], rvaf[2];
    int64_t min_cov;
    int iters;
};

template <class WCtx>
KM_HD Quant2 quant_pair(const WCtx& wctx, const GraphScratch& S, const uint32_t* counts, int n_nodes, const PathView& path,
                        const PathView& range, bool path_first, int lane8, bool allow_jump) {
    const int lane = wctx.tid(), nl = wctx.nt();
    unsigned long long h_p = 0ull, h_r = 0ull, g_pr = 0ull, g_pp = 0ull;
    uint32_t mn = 0xFFFFFFFFu;
    bool twin = false;                 // both columns are the same slice of the reference
    if (path.bub_nk >= 0 || !path.idx) {
        // The identity path or a simple bubble's path (graph.h): which node sits where is known in closed form -- nodes
        // [x0, x1) once in position order, then the chain (novel nodes, each once, none of them in the reference), then
        // nodes [y0, y1) -- so the occurrence counts are interval arithmetic: a node occurs twice exactly when both runs
        // hold it (a tandem duplication), and the only pass left is the one that sums the counts.
        const int e = path.begin + path.len;
        int x0 = path.begin, x1 = e, y0 = 0, y1 = 0;
        if (path.bub_nk >= 0) {
            const int a1 = path.bub_a + 1, tail = a1 + path.bub_nk;          // first position of the chain / after it
            x1 = e < a1 ? e : a1;
            if (x0 > x1) x0 = x1;
            const int s0 = path.begin > tail ? path.begin : tail;
            if (e > s0) { y0 = path.bub_b + (s0 - tail); y1 = path.bub_b + (e - tail); }
        }
        auto overlap = [](int p0, int p1, int q0, int q1) { const int lo = p0 > q0 ? p0 : q0, hi = p1 < q1 ? p1 : q1; return hi > lo ? hi - lo : 0; };
        const int r0 = range.begin, r1 = range.begin + range.len;
        g_pp = (unsigned long long)(path.len + 2 * overlap(x0, x1, y0, y1));
        g_pr = (unsigned long long)(overlap(x0, x1, r0, r1) + overlap(y0, y1, r0, r1));
        // the counts are summed run by run (counts -> float32, PathQuant.py:99: an integer)
        auto take = [&](uint32_t c) { h_p += (unsigned long long)(float)c; mn = c < mn ? c : mn; };
        for (int i = x0 + lane; i < x1; i += nl) take(counts[i]);
        if (path.bub_nk >= 0) {
            const int a1 = path.bub_a + 1;
            int c0 = path.begin - a1, c1 = e - a1;
            c0 = c0 < 0 ? 0 : c0; c1 = c1 > path.bub_nk ? path.bub_nk : c1;
            for (int j = c0 + lane; j < c1; j += nl) take(counts[path.idx[j]]);
            for (int i = y0 + lane; i < y1; i += nl) take(counts[i]);
        }
        const bool same = path.bub_nk < 0 && x0 == r0 && x1 == r1;        // the reference against itself: one sum serves both
        twin = same;
        if (!same) for (int p = lane; p < range.len; p += nl) h_r += (unsigned long long)(float)counts[range.begin + p];
        h_p = warp_sum64(h_p); h_r = same ? h_p : warp_sum64(h_r);
    } else {
        const uint32_t one = 1u << (8 * lane8);
        uint32_t* occ = reinterpret_cast<uint32_t*>(S.occ);
        for (int p = lane; p < path.len; p += nl) {
            const int node = pv_at(path, p);
            const uint32_t c = counts[node];
            atomic_add32(&occ[node], one);
            h_p += (unsigned long long)(float)c;                    // counts -> float32 (PathQuant.py:99), an integer
            g_pr += (unsigned long long)(node >= range.begin && node < range.begin + range.len);
            mn = c < mn ? c : mn;
        }
        for (int p = lane; p < range.len; p += nl) h_r += (unsigned long long)(float)counts[range.begin + p];
        wctx.sync();
        for (int p = lane; p < path.len; p += nl) g_pp += (unsigned long long)((occ[pv_at(path, p)] >> (8 * lane8)) & 255u);
        wctx.sync();
        for (int p = lane; p < path.len; p += nl) atomic_add32(&occ[pv_at(path, p)], 0u - one);
        h_p = warp_sum64(h_p); h_r = warp_sum64(h_r); g_pr = warp_sum64(g_pr); g_pp = warp_sum64(g_pp);
    }
    mn = warp_min32(mn);
    // every lane solves (the sums are broadcast): redundant lanes cost nothing, and a long refinement can use them
    // (refine_jump evaluates its closed form at 32 steps at once)
    h_p = warp_bcast64(h_p); h_r = warp_bcast64(h_r); g_pr = warp_bcast64(g_pr); g_pp = warp_bcast64(g_pp);
    Quant2 q;
    if (twin) {
        // Two identical columns of len ones: G = len * [[1, 1], [1, 1]], h = S * [1, 1].  The minimum-norm solution is
        // S / (2 len) for both; the first refine_coef step finds a zero gradient (to rounding) and stops.  Written down
        // instead of computed: this is the Reference row of every target (whose numbers adjust_for_reference then
        // replaces, PathQuant.py:151-154 -- only "all zero or not" survives), a third of all rows.
        const double c = range.len > 0 ? (double)h_p / (2.0 * (double)range.len) : 0.0;
        q.coef[0] = q.coef[1] = c;
        q.rvaf[0] = q.rvaf[1] = c == 0.0 ? 0.0 : 0.5;
        q.iters = 1;
        q.min_cov = (int64_t)mn;
    } else {
        // exact integer normal equations in the reference's column order
        const unsigned long long g_rr = (unsigned long long)range.len;
        const unsigned long long a00 = path_first ? g_pp : g_rr, a11 = path_first ? g_rr : g_pp, a01 = g_pr;
        double G[4], h[2];
        G[0] = (double)a00; G[1] = G[2] = (double)a01; G[3] = (double)a11;
        h[0] = (double)(path_first ? h_p : h_r); h[1] = (double)(path_first ? h_r : h_p);
        const long long det = (long long)a00 * (long long)a11 - (long long)a01 * (long long)a01;
        double* c = q.coef;
        if (det != 0) {                                          // full rank: Cramer on exact integers
            c[0] = det2(h[0], h[1], G[1], G[3]) / (double)det;
            c[1] = det2(G[0], G[1], h[0], h[1]) / (double)det;
        } else if (a00 + a11 == 0ull) {
            c[0] = c[1] = 0.0;
        } else {
            // rank 1: G = lambda u u^T, lambda = trace; any non-zero column v of G is parallel to u,
            // and the minimum-norm solution (what lstsq returns, PathQuant.py:116) is v (v.h) / (|v|^2 lambda)
            const double v0 = a00 >= a11 ? G[0] : G[1], v1 = a00 >= a11 ? G[1] : G[3];
            const double sc = (v0 * h[0] + v1 * h[1]) / ((v0 * v0 + v1 * v1) * (G[0] + G[3]));
            c[0] = v0 * sc; c[1] = v1 * sc;
        }
        q.iters = refine_and_ratio<2>(lane, nl, G, h, n_nodes, c, q.rvaf, allow_jump);
        q.min_cov = (int64_t)mn;
    }
    return q;
}

// One output row.  `variant` is the (possibly clipped) path, `refv` the (possibly clipped) reference.
KM_HD void write_row(const ResultView& R, const WalkView& W, const uint8_t* kmers, int t, int k, int row_index, int kind,
                     const PathView& refv, const PathView& variant, const Diff& df, int path_id, int offset, int cluster_id,
                     int cluster_n, int iters, int64_t mc, double rvaf, double expr, double ref_rvaf, double ref_expr, int cut_known = -1) {
    int dl, il;
    const int type = classify(kmers, refv, variant, df, &dl, &il, cut_known);
    Row& row = R.rows[row_index];
    row.target = t; row.kind = kind; row.type = type;
    row.name_start = df.start + k + offset; row.name_end = df.end_ref + 1 + offset;
    row.path_id = path_id; row.var_begin = variant.begin; row.var_end = variant.begin + variant.len;
    row.ref_begin = refv.begin; row.ref_end = refv.begin + refv.len;
    row.del_begin = refv.begin + df.start; row.del_len = dl;
    row.ins_begin = variant.begin + df.start; row.ins_len = il;
    row.start_off = offset; row.cluster_id = cluster_id; row.cluster_n = cluster_n; row.n_iter = iters;
    row.min_cov = mc;
    row.rvaf = rvaf; row.expr = expr; row.ref_rvaf = ref_rvaf; row.ref_expr = ref_expr;
    if (iters < 0) atomic_or32(&W.status[t], KM_ST_SOLVER_WATCHDOG);
    if (refv.len - (df.end_ref - df.start) + (df.end_var - df.start) != variant.len)     // MutationFinder.py:431-440
        atomic_or32(&W.status[t], KM_ST_NAME_MISMATCH);
}

// the clipped columns of cluster c (MutationFinder.py:700-723): cols[0] = reference slice,
// cols[1 + j] = member j in join order; members[] receives the path numbers
KM_HD void cluster_columns(const GraphScratch& S, const ResultView& R, const GraphDims& d, int n_paths, int first_path, int c,
                           const int32_t* crec, PathView* cols, int32_t* members) {
    const int lo = crec[4 * c], hi = crec[4 * c + 1], size = crec[4 * c + 2];
    for (int p = 0; p < n_paths; ++p) {
        const int gcode = S.grp[p];
        if (gcode >= 0 && (gcode & 0xFFFF) == c) members[gcode >> 16] = p;
    }
    int span = 0;
    for (int j = 0; j < size; ++j) {
        const int p = members[j];
        int a = S.pdiff[4 * p + 2] - S.pdiff[4 * p + 1] + 1;      // abs(end_var - end_ref + 1) (:710-712)
        a = a < 0 ? -a : a;
        span = a > span ? a : span;
    }
    const int off0 = lo - span > 0 ? lo - span : 0;             // (:713)
    // Python slices clamp to the sequence (:714, :720)
    const int ref_stop = hi < d.L ? hi : d.L;
    cols[0] = range_view(off0, ref_stop - off0 > 0 ? ref_stop - off0 : 0);
    for (int j = 0; j < size; ++j) {
        const int p = members[j];
        PathView v = path_view(S, R, first_path, p);
        const int plen = v.len;
        int stop = S.pdiff[4 * p + 2] + hi - S.pdiff[4 * p + 1];   // (:719)
        stop = stop < plen ? stop : plen;
        const int beg = off0 < plen ? off0 : plen;
        v.begin = beg;
        v.len = stop - beg > 0 ? stop - beg : 0;
        cols[1 + j] = v;
    }
}

// quantify_paths + quantify_clusters for target t.  `dims`, `n_paths`, `first_path` come from
// graph_target, which also reserved rows [first_row, first_row + 2*n_paths).  `sh` = 32 ints of
// CTA-shared memory.  `kmers` = last base of every canonical node, `counts` = their counts (caps are not stored: rows
// never touch them), both in the group's fast memory.  All threads of the group must call this.
// WIDE = false compiles the solver for clusters of several variants out (the bubble pass: such a target goes to the
// general pass instead).
template <class Ctx, bool WIDE = true>
KM_HD void emit_rows_prepared(const Ctx& ctx, const TableView& T, const WalkView& W, const GraphScratch& S,
                              const ResultView& R, int t, const GraphDims& d, int n_paths, int first_path, int first_row, int* sh,
                              const uint8_t* kmers, const uint32_t* counts) {
    const int k = T.k;
    const int tid = ctx.tid();
    const bool allow_jump = !(R.flags & KM_RESULT_NO_REFINE_JUMP);
    const int wid = warp_index(ctx), nw = warp_count(ctx);
    const WarpCtx wctx;
    const int lane = wctx.tid();
    int* slot = sh + 8;                              // CTA-wide reduction scratch
    int* wslot = sh + 16 + wid;                      // this warp's reduction scratch
    const PathView ref = range_view(0, d.L);

    PhaseTimer pt;
    for (int i = tid; i < d.N; i += ctx.nt()) S.occ[i] = 0;     // held nxtF during the tree phase; the solvers need zeros
    // ---- per-path diffs against the whole reference: one warp per path ---------
    for (int p = wid; p < n_paths; p += nw) {
        const PathView alt = path_view(S, R, first_path, p);
        const Diff df = diff_paths(wctx, ref, alt, k, wslot);
        if (lane == 0) {
            S.pdiff[4 * p + 0] = df.start; S.pdiff[4 * p + 1] = df.end_ref;
            S.pdiff[4 * p + 2] = df.end_var; S.pdiff[4 * p + 3] = df.end_ref_overlap;
            S.grp[p] = -2;                       // -2 = still in variant_set
        }
    }
    ctx.sync();

    pt.mark(10);
    // ---- cluster discovery by lane 0 (MutationFinder.py:656-694) ---------------
    // grp[p] = cluster id | (join order << 16), -1 = none; clusters are numbered in seed order
    int32_t* crec = S.grp + S.max_paths;         // lo, hi, size, first row per cluster, in seed order
    if (tid == 0) {
        int n_clusters = 0, n_rows = n_paths;
        for (int seed = 0; seed < n_paths; ++seed) {
            if (S.grp[seed] != -2) continue;     // set.pop() on small ints == ascending order
            int lo = S.pdiff[4 * seed], hi = S.pdiff[4 * seed + 1];
            const int cid = n_clusters;
            S.grp[seed] = cid;
            int size = 1;
            for (;;) {
                int hit = -1;
                for (int v = 0; v < n_paths && hit < 0; ++v) {
                    if (S.grp[v] != -2) continue;
                    const int s = S.pdiff[4 * v], e = S.pdiff[4 * v + 1];
                    if (e >= lo && s <= hi) {
                        if (lo == hi && hi == s && s == e) continue;                    // terminal ITD (:670-671)
                        if (hi == e && (lo == hi || s == e)) continue;                  // quasi-terminal (:672-676)
                        hit = v;
                    }
                }
                if (hit < 0) break;
                S.grp[hit] = cid | (size << 16);
                ++size;
                lo = S.pdiff[4 * hit] < lo ? S.pdiff[4 * hit] : lo;
                hi = S.pdiff[4 * hit + 1] > hi ? S.pdiff[4 * hit + 1] : hi;
            }
            // a lone path equal to the reference forms no cluster (:703-707): the path IS the
            // reference exactly when the common prefix covers both completely
            if (size == 1 && S.ce_len[seed] == d.L && S.pdiff[4 * seed] == d.L) { S.grp[seed] = -1; continue; }
            crec[4 * cid + 0] = lo; crec[4 * cid + 1] = hi; crec[4 * cid + 2] = size; crec[4 * cid + 3] = first_row + n_rows;
            ++n_clusters;
            n_rows += size;
        }
        R.t_n_rows[t] = n_rows;
        R.t_row_first[t] = first_row;
        sh[5] = n_clusters;
    }
    ctx.sync();
    const int n_clusters = sh[5];

    pt.mark(11);
    // ---- two-column jobs, one per warp at a time: the vs_ref row of every path (:613-648) and every
    // single-variant cluster (:758-811) --------------------------------------------------------------
    for (int j = wid; j < n_paths + n_clusters; j += nw) {
        if (j < n_paths) {
            const int p = j;
            const PathView alt = path_view(S, R, first_path, p);
            PhaseTimer jt;
            const Quant2 q = quant_pair(wctx, S, counts, d.N, alt, ref, true, wid, allow_jump);
            jt.mark_warp(51);
            const Diff df = {S.pdiff[4 * p], S.pdiff[4 * p + 1], S.pdiff[4 * p + 2], S.pdiff[4 * p + 3]};
            const int cut = shared_suffix(wctx, kmers, ref, alt, df, wslot);
            if (lane == 0) {
                const bool is_ref = alt.len == d.L && df.start == d.L;      // alt_index == ref_index (:627)
                double c0 = q.coef[0], c1 = q.coef[1], r0 = q.rvaf[0], r1 = q.rvaf[1];
                if (is_ref) {
                    // adjust_for_reference (PathQuant.py:151-154).  With all-zero coef rVAF aliases
                    // coef, so both turn NaN and the `coef >= 0` overwrite skips them.
                    const bool aliased = (c0 > c1 ? c0 : c1) == 0.0;
                    r0 = r1 = NAN;
                    if (aliased) { c0 = c1 = NAN; }
                    else { if (c0 >= 0.0) c0 = -1.0; if (c1 >= 0.0) c1 = -1.0; }   // min(counts) is the cap's -1
                }
                write_row(R, W, kmers, t, k, first_row + p, 0, ref, alt, df, first_path + p, 0, 0, 0, q.iters, q.min_cov, r0, c0, r1, c1, cut);
            }
            jt.mark_warp(52);
        } else {
            const int c = j - n_paths;
            if (crec[4 * c + 2] != 1) continue;                  // wider clusters: whole CTA, below
            // a one-member cluster: columns [ref_clip, alt_clip]
            int p = 0;
            for (int v = 0; v < n_paths; ++v) if (S.grp[v] == c) p = v;      // join order 0 -> code == cluster id
            const int lo = crec[4 * c], hi = crec[4 * c + 1];
            int span = S.pdiff[4 * p + 2] - S.pdiff[4 * p + 1] + 1;          // (:710-712)
            span = span < 0 ? -span : span;
            const int off0 = lo - span > 0 ? lo - span : 0;                  // (:713)
            const int ref_stop = hi < d.L ? hi : d.L;
            const PathView ref_clip = range_view(off0, ref_stop - off0 > 0 ? ref_stop - off0 : 0);
            PathView clip = path_view(S, R, first_path, p);
            const int plen = clip.len;
            int stop = S.pdiff[4 * p + 2] + hi - S.pdiff[4 * p + 1];         // (:719)
            stop = stop < plen ? stop : plen;
            const int beg = off0 < plen ? off0 : plen;
            clip.begin = beg; clip.len = stop - beg > 0 ? stop - beg : 0;
            PhaseTimer jt;
            const Quant2 q = quant_pair(wctx, S, counts, d.N, clip, ref_clip, false, wid, allow_jump);
            jt.mark_warp(48);
            const Diff df = diff_paths(wctx, ref_clip, clip, k, wslot);
            jt.mark_warp(49);
            const int cut = shared_suffix(wctx, kmers, ref_clip, clip, df, wslot);
            if (lane == 0)
                write_row(R, W, kmers, t, k, crec[4 * c + 3], 1, ref_clip, clip, df, first_path + p, off0, c + 1, 1, q.iters,
                          q.min_cov, q.rvaf[1], q.coef[1], q.rvaf[0], q.coef[0], cut);
            jt.mark_warp(50);
        }
    }
    ctx.sync();

    pt.mark(12);
    // ---- clusters of several variants (MutationFinder.py:700-723, 758-811): whole CTA, general solver ----
    double* coef = S.vec + 4 * S.max_cols;
    double* rvaf = S.vec + 5 * S.max_cols;
    PathView* cols = S.cols;
    int32_t* members = S.members;
    for (int c = 0; c < n_clusters; ++c) {
        const int size = crec[4 * c + 2];
        if (size == 1) continue;
        if (!WIDE || size + 1 > S.max_cols) {
            if (tid == 0) {
                const uint32_t before = atomic_or32(&W.status[t], S.retry ? KM_ST_RETRY_LARGE : (uint32_t)KM_ST_TOO_MANY_COLS);
                if (S.retry && !(before & KM_ST_RETRY_LARGE)) defer_to_general(R, W, t);
            }
            continue;          // rows stay unset; the general pass redoes the target, or the host refuses it
        }
        if (!WIDE) continue;
        if (tid == 0) cluster_columns(S, R, d, n_paths, first_path, c, crec, cols, members);
        ctx.sync();
        const int offset = cols[0].begin;
        const PathView ref_clip = cols[0];
        gram_columns(ctx, S, counts, cols, size + 1);
        // The solve is one long dependent chain (solve_wide: warp 0); what each member's row needs besides -- its diff
        // against the clipped reference, its minimum count, the suffix its two strings share -- does not depend on it and
        // is found by the other warps meanwhile, a member each (it used to follow the solve, 38,000 cycles later).
        int32_t* mdiff = reinterpret_cast<int32_t*>(S.vec + 6 * S.max_cols);       // [4 per member]   (free parts of S.vec)
        long long* mmc = reinterpret_cast<long long*>(S.vec + S.max_cols);         // [member]
        int32_t* mcut = reinterpret_cast<int32_t*>(S.vec + 3 * S.max_cols);        // [member]
        auto member_job = [&](int j) {
            const PathView clip = cols[1 + j];
            const Diff df = diff_paths(wctx, ref_clip, clip, k, wslot);
            const int64_t mc = min_count(wctx, counts, clip, wslot);
            const int cut = shared_suffix(wctx, kmers, ref_clip, clip, df, wslot);
            if (lane == 0) {
                mdiff[4 * j] = df.start; mdiff[4 * j + 1] = df.end_ref; mdiff[4 * j + 2] = df.end_var; mdiff[4 * j + 3] = df.end_ref_overlap;
                mmc[j] = mc; mcut[j] = cut;
            }
        };
        if (nw > 1) {
            if (wid == 0) {
                const int it = solve_wide(lane, wctx.nt(), S, size + 1, d.N, coef, rvaf, allow_jump);
                if (lane == 0) sh[4] = it;
            } else {
                for (int j = wid - 1; j < size; j += nw - 1) member_job(j);
            }
        } else {
            const int it = solve_wide(lane, wctx.nt(), S, size + 1, d.N, coef, rvaf, allow_jump);
            if (lane == 0) sh[4] = it;
            for (int j = 0; j < size; ++j) member_job(j);
        }
        ctx.sync();
        const int iters = sh[4];
        for (int j = tid; j < size; j += ctx.nt()) {
            const Diff df = {mdiff[4 * j], mdiff[4 * j + 1], mdiff[4 * j + 2], mdiff[4 * j + 3]};
            write_row(R, W, kmers, t, k, crec[4 * c + 3] + j, 1, ref_clip, cols[1 + j], df, first_path + members[j], offset, c + 1, size,
                      iters, mmc[j], rvaf[1 + j], coef[1 + j], rvaf[0], coef[0], mcut[j]);
        }
        ctx.sync();
    }
    pt.mark(13);
}

// The same for the CTA-per-target passes: last bases and counts are first copied from the canonical node arrays into the
// scratch (shared memory in the small passes), in place of the two distance arrays, which are dead once the paths are
// materialised -- every solver pass reads the counts, naming is a serial scan over last bases.
template <class Ctx>
KM_HD void emit_rows(const Ctx& ctx, const TableView& T, const WalkView& W, const GraphScratch& S,
                     const ResultView& R, int t, const GraphDims& d, int n_paths, int first_path, int first_row, int* sh) {
    const int tid = ctx.tid();
    const int64_t nbase = W.node_off[t];
    uint8_t* last_s = reinterpret_cast<uint8_t*>(S.dist2);
    for (int i = tid; i < d.N - 2; i += ctx.nt()) last_s[i] = (uint8_t)(R.out_kmer[nbase + i] & 3ull);
    uint32_t* cnt_s = reinterpret_cast<uint32_t*>(S.dist);
    for (int i = tid; i < d.N - 2; i += ctx.nt()) cnt_s[i] = R.out_count[nbase + i];
    emit_rows_prepared(ctx, T, W, S, R, t, d, n_paths, first_path, first_row, sh, last_s, cnt_s);
}

}  // namespace km
