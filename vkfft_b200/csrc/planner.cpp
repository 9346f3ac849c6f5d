// Planner: turns a b200fft_desc into an ordered list of kernel launches (see plan.h).
//
// Decisions the reference makes in VkFFTScheduler (vkFFT_Scheduler.h:2223-3299: number of uploads
// :2582-2650, axis split :2651-2893, temp buffer :2902-2944) and VkFFTPlanAxis (vkFFT_Plan_FFT.h:252-417
// strides, :582-645 grid) are made here against the table of ahead-of-time compiled kernels:
//   * an axis whose length has a single-pass kernel  -> one launch (one HBM read + one HBM write);
//   * longer contiguous axes                          -> Four-Step with 2 or 3 launches
//     (strided sub-FFTs + phase multiply, then contiguous sub-FFTs with a transposed, coalesced store so the
//      result is in natural order -- the reference's reorderFourStep=1 behaviour, vkFFT_4step.h:31-119);
//   * axes >= 1                                       -> interleaved-lines ("COLS") kernels.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstdio>
#include <cstring>

#include "plan.h"

namespace b200fft {
namespace {

struct Dim {
    uint64_t n;
    int64_t is, os;  // input / output stride in complex elements
};

int lut_for(PlanGraph& g, const b2_kernel_info* k) {
    std::vector<int> r(k->radices, k->radices + k->ns);
    for (size_t i = 0; i < g.luts.size(); ++i)
        if (g.luts[i].prec == k->prec && g.luts[i].radices == r) return (int)i;
    g.luts.push_back(LutSpec{k->prec, r});
    return (int)g.luts.size() - 1;
}
int tw_for(PlanGraph& g, int prec, uint64_t M) {
    for (size_t i = 0; i < g.tws.size(); ++i)
        if (g.tws[i].prec == prec && g.tws[i].M == M) return (int)i;
    g.tws.push_back(TwSpec{prec, M});
    return (int)g.tws.size() - 1;
}

// merge neighbouring dims that are contiguous in both the input and the output addressing
std::vector<Dim> merge_dims(std::vector<Dim> d) {
    std::vector<Dim> out;
    for (const Dim& x : d) {
        if (x.n == 1) continue;
        if (!out.empty()) {
            Dim& b = out.back();
            if ((int64_t)b.n * b.is == x.is && (int64_t)b.n * b.os == x.os) {
                b.n *= x.n;
                continue;
            }
        }
        out.push_back(x);
    }
    return out;
}

struct PassReq {
    int kind, n, inv, ops;
    int64_t in_es, out_es;
    Dim group;               // lines handled Q at a time by one CTA
    std::vector<Dim> outer;  // remaining line dimensions
    int in_role, out_role;
    uint64_t twM = 0;
    double scale = 1.0;
    const char* what = "";
};

// Emit the launches for one PassReq (more than one only if there are more than B2_MAX_OUTER outer dims).
int emit(PlanGraph& g, std::vector<PassPlan>& list, const PassReq& rq) {
    const b2_kernel_info* k = b2_find_kernel(rq.kind, g.prec, rq.n, rq.inv, rq.ops & B2_OP_TWIDDLE_OUT);
    if (!k) return R_UNSUPPORTED_FFT_LENGTH;
    std::vector<Dim> outer = merge_dims(rq.outer);
    // peel outermost dims into separate launches until at most B2_MAX_OUTER remain
    std::vector<Dim> peeled;
    while (outer.size() > B2_MAX_OUTER) {
        peeled.push_back(outer.back());
        outer.pop_back();
    }
    uint64_t npeel = 1;
    for (const Dim& p : peeled) npeel *= p.n;
    for (uint64_t pi = 0; pi < npeel; ++pi) {
        int64_t ioff = 0, ooff = 0;
        uint64_t rem = pi;
        for (const Dim& p : peeled) {
            uint64_t c = rem % p.n;
            rem /= p.n;
            ioff += (int64_t)c * p.is;
            ooff += (int64_t)c * p.os;
        }
        PassPlan pp;
        pp.k = k;
        b2_pass_params& P = pp.P;
        P.in_es = rq.in_es; P.out_es = rq.out_es;
        P.in_gs = rq.group.is; P.out_gs = rq.group.os;
        P.G = (uint32_t)rq.group.n;
        uint64_t grid = (rq.group.n + k->q - 1) / k->q;
        for (int d = 0; d < B2_MAX_OUTER; ++d) {
            if (d < (int)outer.size()) {
                P.nb[d] = (uint32_t)outer[d].n; P.in_bs[d] = outer[d].is; P.out_bs[d] = outer[d].os;
            } else {
                P.nb[d] = 1; P.in_bs[d] = 0; P.out_bs[d] = 0;
            }
            grid *= P.nb[d];
        }
        if (grid == 0 || grid > 0x7fffffffull) return R_UNSUPPORTED_FFT_LENGTH;
        pp.grid = (unsigned)grid;
        P.n = rq.n;
        P.ops = rq.ops;
        P.inverse = rq.inv;
        P.scale = rq.scale;
        pp.in_role = rq.in_role; pp.out_role = rq.out_role;
        pp.in_off = ioff; pp.out_off = ooff;
        pp.lut_id = lut_for(g, k);
        if (rq.ops & B2_OP_TWIDDLE_OUT) pp.tw_id = tw_for(g, g.prec, rq.twM);
        char buf[256];
        snprintf(buf, sizeof buf, "%s n=%d %s grid=%u threads=%d smem=%d  %s", rq.what, rq.n, k->name, pp.grid,
                 k->threads, k->smem_bytes, rq.inv ? "inverse" : "forward");
        pp.note = buf;
        list.push_back(pp);
    }
    return R_SUCCESS;
}

bool have(const PlanGraph& g, int kind, uint64_t n, int ops) {
    if (n > 0x7fffffffull) return false;
    return b2_find_kernel(kind, g.prec, (int)n, 0, ops) != nullptr;
}

// Factor N for Four-Step.  All factors but the last run as interleaved-line passes with the phase multiply,
// the last one runs on contiguous lines with a transposed store.  Returns empty if impossible.
std::vector<uint64_t> split_four_step(const PlanGraph& g, uint64_t N) {
    // test hook: B200FFT_FOUR_STEP_SPLIT="n1,n2[,n3]" forces a factorisation (used by the CPU tests to reach
    // the three-pass code with small transforms)
    if (const char* e = getenv("B200FFT_FOUR_STEP_SPLIT")) {
        std::vector<uint64_t> f;
        uint64_t prod = 1;
        for (const char* p = e; *p;) {
            char* end;
            uint64_t v = strtoull(p, &end, 10);
            if (end == p) break;
            f.push_back(v); prod *= v;
            p = (*end == ',') ? end + 1 : end;
        }
        if (prod == N && f.size() >= 2 && f.size() <= 3) return f;
    }
    std::vector<uint64_t> best;
    uint64_t best_cost = ~0ull;
    // two factors
    for (uint64_t n2 = 2; n2 * 2 <= N; ++n2) {
        if (N % n2) continue;
        uint64_t n1 = N / n2;
        if (!have(g, B2_KIND_ROWS_TOUT, n2, 0) || !have(g, B2_KIND_COLS, n1, B2_OP_TWIDDLE_OUT)) continue;
        // prefer balanced factors, the contiguous one not smaller than the strided one
        uint64_t cost = std::max(n1, n2) * 4 + (n2 < n1 ? 2 : 0);
        if (cost < best_cost) { best_cost = cost; best = {n1, n2}; }
    }
    const uint64_t two_level_limit = (g.prec == B2_PREC_F32) ? (1ull << 22) : (1ull << 21);
    if (!best.empty() && N <= two_level_limit) return best;
    // three factors
    std::vector<uint64_t> best3;
    uint64_t best3_cost = ~0ull;
    for (uint64_t n3 = 2; n3 * 4 <= N; ++n3) {
        if (N % n3 || !have(g, B2_KIND_ROWS_TOUT, n3, 0)) continue;
        uint64_t rest = N / n3;
        for (uint64_t n2 = 2; n2 * 2 <= rest; ++n2) {
            if (rest % n2 || !have(g, B2_KIND_COLS, n2, B2_OP_TWIDDLE_OUT)) continue;
            uint64_t n1 = rest / n2;
            if (!have(g, B2_KIND_COLS, n1, B2_OP_TWIDDLE_OUT)) continue;
            uint64_t cost = std::max(n1, std::max(n2, n3)) * 4 + (n3 < n1 ? 1 : 0) + (n3 < n2 ? 1 : 0);
            if (cost < best3_cost) { best3_cost = cost; best3 = {n1, n2, n3}; }
        }
    }
    if (!best3.empty()) return best3;
    return best;
}

struct AxisIO {
    int in_role, out_role;                 // where this axis reads / writes
    const uint64_t* in_stride;             // per-dimension strides (elements) on the input side
    const uint64_t* out_stride;
    uint64_t in_batch_stride, out_batch_stride;
    double scale;                          // 1.0 or the normalisation factor (applied by the last launch)
};

// contiguous axis 0
int plan_axis0(PlanGraph& g, std::vector<PassPlan>& list, int inv, const AxisIO& io) {
    const b200fft_desc& d = g.desc;
    const uint64_t N = d.size[0];
    std::vector<Dim> lines;  // all dims except axis 0
    for (uint32_t a = 1; a < d.fft_dim; ++a)
        lines.push_back(Dim{d.size[a], (int64_t)io.in_stride[a - 1], (int64_t)io.out_stride[a - 1]});
    lines.push_back(Dim{g.batches, (int64_t)io.in_batch_stride, (int64_t)io.out_batch_stride});
    const int sc_ops = (io.scale != 1.0) ? B2_OP_SCALE : 0;

    uint64_t max_single = ~0ull;
    if (const char* e = getenv("B200FFT_MAX_SINGLE_PASS")) max_single = strtoull(e, nullptr, 10);
    if (N <= max_single && have(g, B2_KIND_ROWS, N, 0)) {
        std::vector<Dim> m = merge_dims(lines);
        PassReq rq{};
        rq.kind = B2_KIND_ROWS; rq.n = (int)N; rq.inv = inv; rq.ops = sc_ops;
        rq.in_es = 1; rq.out_es = 1;
        if (m.empty()) rq.group = Dim{1, (int64_t)N, (int64_t)N};
        else { rq.group = m[0]; m.erase(m.begin()); }
        rq.outer = m;
        rq.in_role = io.in_role; rq.out_role = io.out_role;
        rq.scale = io.scale;
        rq.what = "axis0 single-pass";
        return emit(g, list, rq);
    }
    std::vector<uint64_t> f = split_four_step(g, N);
    if (f.empty()) return R_UNSUPPORTED_FFT_LENGTH;
    g.temp_elems = std::max<uint64_t>(g.temp_elems, g.batch_stride * g.batches);
    // sequences on the temp buffer use the main buffer's layout
    std::vector<Dim> seq_in_to_tmp, seq_tmp_to_tmp, seq_tmp_to_out, seq_in_to_in;
    for (uint32_t a = 1; a < d.fft_dim; ++a) {
        seq_in_to_tmp.push_back(Dim{d.size[a], (int64_t)io.in_stride[a - 1], (int64_t)g.stride[a - 1]});
        seq_tmp_to_tmp.push_back(Dim{d.size[a], (int64_t)g.stride[a - 1], (int64_t)g.stride[a - 1]});
        seq_tmp_to_out.push_back(Dim{d.size[a], (int64_t)g.stride[a - 1], (int64_t)io.out_stride[a - 1]});
        seq_in_to_in.push_back(Dim{d.size[a], (int64_t)io.in_stride[a - 1], (int64_t)io.in_stride[a - 1]});
    }
    seq_in_to_tmp.push_back(Dim{g.batches, (int64_t)io.in_batch_stride, (int64_t)g.batch_stride});
    seq_tmp_to_tmp.push_back(Dim{g.batches, (int64_t)g.batch_stride, (int64_t)g.batch_stride});
    seq_tmp_to_out.push_back(Dim{g.batches, (int64_t)g.batch_stride, (int64_t)io.out_batch_stride});
    seq_in_to_in.push_back(Dim{g.batches, (int64_t)io.in_batch_stride, (int64_t)io.in_batch_stride});

    int rc;
    if (f.size() == 2) {
        const uint64_t N1 = f[0], N2 = f[1];
        PassReq a{};
        a.kind = B2_KIND_COLS; a.n = (int)N1; a.inv = inv; a.ops = B2_OP_TWIDDLE_OUT;
        a.in_es = (int64_t)N2; a.out_es = (int64_t)N2;
        a.group = Dim{N2, 1, 1};
        a.outer = seq_in_to_tmp;
        a.in_role = io.in_role; a.out_role = ROLE_TEMP;
        a.twM = N;
        a.what = "four-step 1/2 strided+phase";
        if ((rc = emit(g, list, a)) != R_SUCCESS) return rc;
        PassReq b{};
        b.kind = B2_KIND_ROWS_TOUT; b.n = (int)N2; b.inv = inv; b.ops = sc_ops;
        b.in_es = 1; b.out_es = (int64_t)N1;
        b.group = Dim{N1, (int64_t)N2, 1};
        b.outer = seq_tmp_to_out;
        b.in_role = ROLE_TEMP; b.out_role = io.out_role;
        b.scale = io.scale;
        b.what = "four-step 2/2 contiguous+transpose";
        return emit(g, list, b);
    }
    const uint64_t N1 = f[0], N2 = f[1], N3 = f[2], M = N2 * N3;
    PassReq a{};
    a.kind = B2_KIND_COLS; a.n = (int)N1; a.inv = inv; a.ops = B2_OP_TWIDDLE_OUT;
    a.in_es = (int64_t)M; a.out_es = (int64_t)M;
    a.group = Dim{M, 1, 1};
    // first pass runs in place on its input when that is the main buffer, otherwise it moves to temp
    const bool a_inplace = (io.in_role == ROLE_BUFFER);
    a.outer = a_inplace ? seq_in_to_in : seq_in_to_tmp;
    a.in_role = io.in_role; a.out_role = a_inplace ? io.in_role : ROLE_TEMP;
    a.twM = N;
    a.what = "four-step 1/3 strided+phase";
    if ((rc = emit(g, list, a)) != R_SUCCESS) return rc;
    PassReq b{};
    b.kind = B2_KIND_COLS; b.n = (int)N2; b.inv = inv; b.ops = B2_OP_TWIDDLE_OUT;
    b.in_es = (int64_t)N3; b.out_es = (int64_t)N3;
    b.group = Dim{N3, 1, 1};
    b.outer.push_back(Dim{N1, (int64_t)M, (int64_t)M});
    {
        const std::vector<Dim>& s = a_inplace ? seq_in_to_tmp : seq_tmp_to_tmp;
        b.outer.insert(b.outer.end(), s.begin(), s.end());
    }
    b.in_role = a.out_role; b.out_role = ROLE_TEMP;
    b.twM = M;
    b.what = "four-step 2/3 strided+phase";
    if ((rc = emit(g, list, b)) != R_SUCCESS) return rc;
    PassReq c{};
    c.kind = B2_KIND_ROWS_TOUT; c.n = (int)N3; c.inv = inv; c.ops = sc_ops;
    c.in_es = 1; c.out_es = (int64_t)(N1 * N2);
    c.group = Dim{N1, (int64_t)M, 1};
    c.outer.push_back(Dim{N2, (int64_t)N3, (int64_t)N1});
    c.outer.insert(c.outer.end(), seq_tmp_to_out.begin(), seq_tmp_to_out.end());
    c.in_role = ROLE_TEMP; c.out_role = io.out_role;
    c.scale = io.scale;
    c.what = "four-step 3/3 contiguous+transpose";
    return emit(g, list, c);
}

// strided axis a >= 1
int plan_axis_strided(PlanGraph& g, std::vector<PassPlan>& list, uint32_t axis, int inv, const AxisIO& io) {
    const b200fft_desc& d = g.desc;
    const uint64_t N = d.size[axis];
    if (!have(g, B2_KIND_COLS, N, 0)) return R_UNSUPPORTED_FFT_LENGTH;
    std::vector<Dim> lines;
    lines.push_back(Dim{d.size[0], 1, 1});
    for (uint32_t a = 1; a < d.fft_dim; ++a) {
        if (a == axis) continue;
        lines.push_back(Dim{d.size[a], (int64_t)io.in_stride[a - 1], (int64_t)io.out_stride[a - 1]});
    }
    lines.push_back(Dim{g.batches, (int64_t)io.in_batch_stride, (int64_t)io.out_batch_stride});
    // keep the unit-stride dimension first even if it has extent 1
    std::vector<Dim> m = merge_dims(lines);
    PassReq rq{};
    rq.kind = B2_KIND_COLS; rq.n = (int)N; rq.inv = inv; rq.ops = (io.scale != 1.0) ? B2_OP_SCALE : 0;
    rq.in_es = (int64_t)io.in_stride[axis - 1]; rq.out_es = (int64_t)io.out_stride[axis - 1];
    if (!m.empty() && m[0].is == 1 && m[0].os == 1) { rq.group = m[0]; m.erase(m.begin()); }
    else rq.group = Dim{1, 1, 1};
    rq.outer = m;
    rq.in_role = io.in_role; rq.out_role = io.out_role;
    rq.scale = io.scale;
    rq.what = "strided axis";
    return emit(g, list, rq);
}

int plan_direction(PlanGraph& g, std::vector<PassPlan>& list, int inv) {
    const b200fft_desc& d = g.desc;
    // order of axes: forward 0,1,2..; inverse ..2,1,0 (vkFFT_RunApp.h:111-321 / :466-651)
    std::vector<uint32_t> axes;
    for (uint32_t a = 0; a < d.fft_dim; ++a)
        if (!d.omit_dimension[a]) axes.push_back(a);
    if (inv) std::reverse(axes.begin(), axes.end());
    double norm = 1.0;
    if (inv && d.normalize)
        for (uint32_t a : axes) norm /= (double)d.size[a];
    // out-of-place plumbing (documentation/VkFFT_API_guide.tex:365-376): the first launch reads the formatted
    // input, the last launch writes the formatted output, everything in between lives in `buffer`.
    const bool fmt_in = d.is_input_formatted != 0, fmt_out = d.is_output_formatted != 0;
    for (size_t i = 0; i < axes.size(); ++i) {
        const bool first = (i == 0), last = (i + 1 == axes.size());
        AxisIO io{};
        io.in_role = ROLE_BUFFER; io.out_role = ROLE_BUFFER;
        io.in_stride = g.stride; io.out_stride = g.stride;
        io.in_batch_stride = g.batch_stride; io.out_batch_stride = g.batch_stride;
        static thread_local uint64_t istr[B200FFT_MAX_DIMS], ostr[B200FFT_MAX_DIMS];
        if (!inv) {
            if (first && fmt_in) {
                io.in_role = ROLE_INPUT;
                for (int k = 0; k < B200FFT_MAX_DIMS; ++k) istr[k] = d.input_stride[k];
                io.in_stride = istr; io.in_batch_stride = d.input_stride[d.fft_dim - 1] ;
            }
            if (last && fmt_out) {
                io.out_role = ROLE_OUTPUT;
                for (int k = 0; k < B200FFT_MAX_DIMS; ++k) ostr[k] = d.output_stride[k];
                io.out_stride = ostr; io.out_batch_stride = d.output_stride[d.fft_dim - 1];
            }
        } else {
            // inverse mirrors the forward data flow: it consumes what forward produced and returns it to where
            // forward read from (outputBuffer -> ... -> buffer, or -> inputBuffer with inverseReturnToInputBuffer)
            if (first && fmt_out) {
                io.in_role = ROLE_OUTPUT;
                for (int k = 0; k < B200FFT_MAX_DIMS; ++k) istr[k] = d.output_stride[k];
                io.in_stride = istr; io.in_batch_stride = d.output_stride[d.fft_dim - 1];
            }
            if (last && fmt_in && d.inverse_return_to_input) {
                io.out_role = ROLE_INPUT;
                for (int k = 0; k < B200FFT_MAX_DIMS; ++k) ostr[k] = d.input_stride[k];
                io.out_stride = ostr; io.out_batch_stride = d.input_stride[d.fft_dim - 1];
            }
        }
        io.scale = last ? norm : 1.0;
        int rc = (axes[i] == 0) ? plan_axis0(g, list, inv, io) : plan_axis_strided(g, list, axes[i], inv, io);
        if (rc != R_SUCCESS) return rc;
    }
    return R_SUCCESS;
}

}  // namespace

int build_plan(const b200fft_desc& din, PlanGraph& g) {
    g = PlanGraph{};
    b200fft_desc d = din;
    if (d.fft_dim == 0) return R_EMPTY_FFTDIM;
    if (d.fft_dim > B200FFT_MAX_DIMS) return R_FFTDIM_GT_MAX;
    if (d.size[0] == 0) return R_EMPTY_SIZE;
    for (uint32_t a = 1; a < B200FFT_MAX_DIMS; ++a)
        if (d.size[a] == 0 || a >= d.fft_dim) d.size[a] = 1;
    if (d.number_batches == 0) d.number_batches = 1;
    if (d.coordinate_features == 0) d.coordinate_features = 1;
    if (d.precision > B200FFT_F64) return R_UNSUPPORTED_FFT_LENGTH;
    if (d.perform_r2c || d.perform_dct || d.perform_dst) return R_UNSUPPORTED_FFT_LENGTH_R2C;
    if (d.omit_dimension[0] && d.perform_r2c) return R_UNSUPPORTED_FFT_OMIT;
    // default strides (vkFFT_InitializeApp.h:994-1040)
    auto fill = [&](uint64_t* s) {
        if (s[0] == 0) s[0] = d.size[0];
        for (int a = 1; a < B200FFT_MAX_DIMS; ++a)
            if (s[a] == 0) s[a] = s[a - 1] * d.size[a];
    };
    fill(d.buffer_stride); fill(d.input_stride); fill(d.output_stride);
    g.desc = d;
    g.prec = (int)d.precision;
    for (int a = 0; a < B200FFT_MAX_DIMS; ++a) g.stride[a] = d.buffer_stride[a];
    g.batches = d.number_batches * d.coordinate_features;
    g.batch_stride = d.buffer_stride[d.fft_dim - 1];
    g.total_elems = g.batches;
    for (uint32_t a = 0; a < d.fft_dim; ++a) g.total_elems *= d.size[a];

    uint32_t naxes = 0;
    g.flops = 0;
    for (uint32_t a = 0; a < d.fft_dim; ++a) {
        if (d.omit_dimension[a]) continue;
        ++naxes;
        double n = (double)d.size[a], l2 = 0;
        for (double t = n; t > 1; t /= 2) l2 += 1;  // exact for powers of two
        if ((d.size[a] & (d.size[a] - 1)) != 0) l2 = std::log2(n);
        g.flops += 5.0 * (double)g.total_elems * l2;
    }
    const uint64_t esz = (g.prec == B2_PREC_F64) ? 16 : 8;
    g.algorithmic_bytes = 2 * esz * g.total_elems * naxes;

    g.has_fwd = !d.make_inverse_plan_only;
    g.has_inv = !d.make_forward_plan_only;
    int rc;
    if (g.has_fwd && (rc = plan_direction(g, g.fwd, 0)) != R_SUCCESS) return rc;
    if (g.has_inv && (rc = plan_direction(g, g.inv, 1)) != R_SUCCESS) return rc;
    if (d.user_temp_buffer && g.temp_elems * esz > d.temp_buffer_size && d.temp_buffer_size != 0)
        return R_USER_TEMP_TOO_SMALL;
    return R_SUCCESS;
}

}  // namespace b200fft
