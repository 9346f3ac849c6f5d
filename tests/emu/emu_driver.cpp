// TEST INFRASTRUCTURE ONLY: runs the engine's kernel bodies on the CPU emulation (see cuda_emu.h).
// Built by tests/emu/build.sh into tests/emu/_build/libb200fft_emu.so and driven from pytest via ctypes.
#include "kernel_inst.cuh"
#include "lut.h"
#include "plan.h"

#include <cstdio>
#include <vector>

using namespace b200fft;

static GenericRegistrar<float, 8> b2_generic_f32_8("generic<float,r<=8>");
static GenericRegistrar<float, 11> b2_generic_f32_11("generic<float,r<=11>");
static GenericRegistrar<float, 16> b2_generic_f32_16("generic<float,r<=16>");
static GenericRegistrar<double, 8> b2_generic_f64_8("generic<double,r<=8>");
static GenericRegistrar<double, 11> b2_generic_f64_11("generic<double,r<=11>");
static GenericRegistrar<double, 16> b2_generic_f64_16("generic<double,r<=16>");
static ElementwiseRegistrar<float> b2_ew_f32("elementwise<float>");
static ElementwiseRegistrar<double> b2_ew_f64("elementwise<double>");

extern "C" int emu_kernel_count() { return b2_kernel_count(); }
extern "C" int emu_kernel_info(int i, int* out /*kind,prec,n,inv,ops,threads,q,tpl,v,smem,ns,r0..r7*/) {
    const b2_kernel_info* k = b2_kernel_at(i);
    int v[] = {k->kind, k->prec, k->n, k->inv, k->ops, k->threads, k->q, k->tpl, k->v, k->smem_bytes, k->ns};
    for (int j = 0; j < 11; ++j) out[j] = v[j];
    for (int j = 0; j < 8; ++j) out[11 + j] = k->radices[j];
    out[19] = k->variant;
    out[20] = k->pipelined;
    return 0;
}

// Run one pass.  `in`/`out` are host arrays of complex T.  Returns 0, or -1 if no such kernel.
extern "C" int emu_run_pass(int kind, int prec, int n, int inv, int ops, int variant, const void* in, void* out, unsigned G,
                            const unsigned* nb, long long in_es, long long out_es, long long in_gs,
                            long long out_gs, const long long* in_bs, const long long* out_bs,
                            unsigned long long twM, unsigned tw_line0, double scale, int log, double* report) {
    const b2_kernel_info* k = b2_find_kernel_variant(kind, prec, n, inv, ops & ~B2_OP_SCALE, variant);
    if (!k) return -1;
    b2_pass_params P{};
    P.in = in; P.out = out;
    std::vector<float> lutf, hif, lof, tilef;
    std::vector<double> lutd, hid, lod, tiled;
    if (prec == B2_PREC_F32) { lutf = make_stage_lut<float>(k->radices, k->ns); P.lut = lutf.data(); }
    else { lutd = make_stage_lut<double>(k->radices, k->ns); P.lut = lutd.data(); }
    if (ops & B2_OP_TWIDDLE_OUT) {
        if (prec == B2_PREC_F32) { make_twolevel<float>(twM, P.tw_shift, hif, lof); P.tw_hi = hif.data(); P.tw_lo = lof.data(); }
        else { make_twolevel<double>(twM, P.tw_shift, hid, lod); P.tw_hi = hid.data(); P.tw_lo = lod.data(); }
    }
    P.in_es = in_es; P.out_es = out_es; P.in_gs = in_gs; P.out_gs = out_gs;
    unsigned grid = (G + k->q - 1) / k->q;
    for (int d = 0; d < B2_MAX_OUTER; ++d) {
        P.nb[d] = nb[d]; P.in_bs[d] = in_bs[d]; P.out_bs[d] = out_bs[d];
        grid *= nb[d];
    }
    P.G = G; P.n = n; P.tw_line0 = tw_line0; P.ops = ops; P.inverse = inv; P.scale = scale;
    b2emu::st().log = log != 0;
    int rc = k->launch(&P, grid, nullptr);
    if (log && report) {
        b2emu::ConflictReport r = b2emu::analyse(k->threads);
        report[0] = r.worst; report[1] = r.mean; report[2] = (double)r.accesses;
    }
    b2emu::st().log = false;
    return rc;
}

// Execute a whole plan (planner.cpp + emulated kernels) on host memory.  bufs[ROLE_*] are host pointers;
// the temp buffer is allocated here.  Returns the planner's VkFFTResult code.
static void* g_emu_kernel = nullptr;
extern "C" void emu_set_kernel(void* k) { g_emu_kernel = k; }

static int emu_run_one(PlanGraph& g, PassPlan& pp, void* const* base) {
    const size_t esz = g.prec == B2_PREC_F64 ? 16 : 8;
    b2_pass_params P = pp.P;
    std::vector<float> lutf, hif, lof;
    std::vector<double> lutd, hid, lod;
    const LutSpec& ls = g.luts[pp.lut_id];
    if (g.prec == B2_PREC_F32) { lutf = make_stage_lut<float>(ls.radices.data(), (int)ls.radices.size()); P.lut = lutf.data(); }
    else { lutd = make_stage_lut<double>(ls.radices.data(), (int)ls.radices.size()); P.lut = lutd.data(); }
    if (pp.tw_id >= 0) {
        uint64_t M = g.tws[pp.tw_id].M;
        if (g.prec == B2_PREC_F32) { make_twolevel<float>(M, P.tw_shift, hif, lof); P.tw_hi = hif.data(); P.tw_lo = lof.data(); }
        else { make_twolevel<double>(M, P.tw_shift, hid, lod); P.tw_hi = hid.data(); P.tw_lo = lod.data(); }
    }
    std::vector<float> a0f, a1f; std::vector<double> a0d, a1d;
    if (pp.aux0_id >= 0) {
        const AuxSpec& a = g.auxs[pp.aux0_id];
        if (g.prec == B2_PREC_F32) { a0f = make_aux<float>(a.kind, a.a, a.b); P.aux0 = a0f.data(); }
        else { a0d = make_aux<double>(a.kind, a.a, a.b); P.aux0 = a0d.data(); }
    }
    if (pp.aux1_id >= 0) {
        const AuxSpec& a = g.auxs[pp.aux1_id];
        if (g.prec == B2_PREC_F32) { a1f = make_aux<float>(a.kind, a.a, a.b); P.aux1 = a1f.data(); }
        else { a1d = make_aux<double>(a.kind, a.a, a.b); P.aux1 = a1d.data(); }
    }
    if (pp.aux0_role == ROLE_KERNEL) P.aux0 = g_emu_kernel;
    const size_t esz_in = g.role_half[pp.in_role] ? 4 : esz, esz_out = g.role_half[pp.out_role] ? 4 : esz;   // half-precision storage
    P.in = (const unsigned char*)base[pp.in_role] + pp.in_off * (pp.in_scalar ? esz_in / 2 : esz_in);
    P.out = (unsigned char*)base[pp.out_role] + pp.out_off * (pp.out_scalar ? esz_out / 2 : esz_out);
    return pp.k->launch(&P, pp.grid, nullptr) ? 4039 : 0;
}

// both launches of a fused Four-Step pair through the fused kernel body (tables built per pass as above)
static int emu_run_fused(PlanGraph& g, PassPlan& pa, PassPlan& pb, void* const* base) {
    const size_t esz = g.prec == B2_PREC_F64 ? 16 : 8;
    b2_fused_params F{};
    std::vector<float> lutf[2], hif, lof;
    std::vector<double> lutd[2], hid, lod;
    PassPlan* pps[2] = {&pa, &pb};
    b2_pass_params* Ps[2] = {&F.A, &F.B};
    for (int i = 0; i < 2; ++i) {
        PassPlan& pp = *pps[i];
        b2_pass_params& P = *Ps[i];
        P = pp.P;
        const LutSpec& ls = g.luts[pp.lut_id];
        if (g.prec == B2_PREC_F32) { lutf[i] = make_stage_lut<float>(ls.radices.data(), (int)ls.radices.size()); P.lut = lutf[i].data(); }
        else { lutd[i] = make_stage_lut<double>(ls.radices.data(), (int)ls.radices.size()); P.lut = lutd[i].data(); }
        if (pp.tw_id >= 0) {
            uint64_t M = g.tws[pp.tw_id].M;
            if (g.prec == B2_PREC_F32) { make_twolevel<float>(M, P.tw_shift, hif, lof); P.tw_hi = hif.data(); P.tw_lo = lof.data(); }
            else { make_twolevel<double>(M, P.tw_shift, hid, lod); P.tw_hi = hid.data(); P.tw_lo = lod.data(); }
        }
        P.in = (const unsigned char*)base[pp.in_role] + pp.in_off * esz;
        P.out = (unsigned char*)base[pp.out_role] + pp.out_off * esz;
    }
    std::vector<uint32_t> ctl(B2_FCTL_WORDS + 2 * (size_t)pa.fz_NU);
    F.ctl = ctl.data();
    F.nseq = pa.fz_nseq; F.U = pa.fz_U; F.NU = pa.fz_NU; F.R = pa.fz_R; F.TA = pa.fz_TA; F.TB = pa.fz_TB; F.reserved = pa.fz_L;
    return pa.fused->launch(&F, 0, nullptr) ? 4039 : 0;
}

extern "C" int emu_exec_plan(const b200fft_desc* d, int inverse, void* buffer, void* input, void* output,
                             int* npasses) {
    PlanGraph g;
    int rc = build_plan(*d, g);
    if (rc != 0) return rc;
    if (inverse == 1 && !g.has_inv) return R_ONLY_FORWARD;      // same checks as b200fft_exec
    if (inverse != 1 && !g.has_fwd) return R_ONLY_INVERSE;
    const size_t esz = g.prec == B2_PREC_F64 ? 16 : 8;
    std::vector<unsigned char> temp(g.temp_elems * esz + 16);
    void* base[ROLE_COUNT] = {buffer, temp.data(), input, output, g_emu_kernel};
    std::vector<PassPlan>& list = (inverse == 1) ? g.inv : g.fwd;
    if (npasses) *npasses = (int)list.size();
    for (size_t i = 0; i < list.size(); ++i) {
        if (list[i].fused && i + 1 < list.size()) {
            if ((rc = emu_run_fused(g, list[i], list[i + 1], base)) != 0) return rc;
            ++i;
            continue;
        }
        if ((rc = emu_run_one(g, list[i], base)) != 0) return rc;
    }
    return 0;
}

// One launch of a plan on caller-supplied buffer + temp (distributed plans: the test plays every rank's launches in
// an order the barriers allow, on two host arrays standing for the two peer windows).
// sync_before[i] (when non-null, capacity 16) reports which launches are preceded by a barrier.
extern "C" int emu_exec_plan_pass(const b200fft_desc* d, int inverse, void* buffer, void* temp, int pass, int* npasses,
                                  int* sync_before) {
    PlanGraph g;
    int rc = build_plan(*d, g);
    if (rc != 0) return rc;
    void* base[ROLE_COUNT] = {buffer, temp, nullptr, nullptr, g_emu_kernel};
    std::vector<PassPlan>& list = (inverse == 1) ? g.inv : g.fwd;
    if (npasses) *npasses = (int)list.size();
    if (sync_before)
        for (size_t i = 0; i < list.size() && i < 16; ++i) sync_before[i] = list[i].sync_before ? 1 : 0;
    if (pass < 0) return 0;   // query only
    if (pass >= (int)list.size()) return 4;
    return emu_run_one(g, list[pass], base);
}

// plan listing without a GPU (same text as b200fft_plan_describe)
extern "C" int emu_describe(const b200fft_desc* d, int inverse, char* dst, int cap) {
    PlanGraph g;
    int rc = build_plan(*d, g);
    if (rc != 0) return rc;
    std::string s;
    static const char* role[] = {"buffer", "temp", "input", "output", "kernel"};
    for (const PassPlan& pp : (inverse == 1 ? g.inv : g.fwd))
        s += pp.note + "  " + role[pp.in_role] + " -> " + role[pp.out_role] + (pp.sync_before ? "  [barrier before]" : "") + "\n";
    snprintf(dst, cap, "%s", s.c_str());
    return 0;
}

// self test of the emulation's checkers: mode 0 = correct exchange through shared memory (own slot, barrier, neighbour's slot),
// 1 = the barrier is missing (read-after-write hazard), 2 = two threads write one word, 3 = write beyond the allocation.
// Returns hazards (modes 0-2) or the out-of-bounds flag (mode 3).
extern "C" int emu_selftest_checkers(int mode) {
    b2emu::State& s = b2emu::st();
    s.hazards = 0; s.smem_oob = false;
    b2emu::launch(2, 64, 64 * sizeof(float), [&](unsigned char* raw) {
        float* sm = reinterpret_cast<float*>(raw);
        const unsigned t = threadIdx.x;
        if (mode == 2) { B2_SMEM_ST(sm, t / 2, 1.0f); return; }
        if (mode == 3) { if (t == 0) B2_SMEM_ST(sm, 64, 1.0f); return; }
        B2_SMEM_ST(sm, t, (float)t);
        if (mode == 0) __syncthreads();
        volatile float v = B2_SMEM_LD(sm, (t + 1) % 64);
        (void)v;
        if (mode == 0) __syncthreads();
    }, false);
    const int r = mode == 3 ? (s.smem_oob ? 1 : 0) : s.hazards.load();
    s.hazards = 0; s.smem_oob = false;
    return r;
}

// host versions of the half <-> float conversions the emulated half-storage kernels use (cplx.cuh); the device uses cvt.f32.f16 /
// cvt.rn.f16x2.f32 -- both are IEEE round-to-nearest-even, tests/test_half_storage.py pins the host ones to numpy
extern "C" unsigned short emu_float_to_half_bits(float f) { return b200fft::b2_float_to_half_bits(f); }
extern "C" float emu_half_bits_to_float(unsigned short h) { return b200fft::b2_half_bits_to_float(h); }
