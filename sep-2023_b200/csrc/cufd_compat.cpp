// cufd_compat.cpp -- sepfwi_cufd: drop-in for the reference's `extern "C" cufd` shot driver
// (DAS_Waveform_Inversion/Ops/FWI/Src/libCUFD.cu:32-820) on top of the handle API.
//
// Same side channels as the reference: para_file.json (Parameter.cpp:17-178),
// survey_file.json (Src_Rec.cu:20-282), Shot_{pr,vx,vz,ett}{id}.bin raw float32
// [nrec][nSteps] in data_dir_name (libCUFD.cu:215-223,755-769).  Handles are cached per
// (device, parameter set) so an L-BFGS loop pays the allocations once, not per evaluation.
#include <cuda_runtime.h>
#include <ctype.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/sepfwi.h"

namespace {

// ---- a very small JSON reader (objects, arrays, numbers, strings, true/false/null) ----
struct JVal {
    enum Kind { Null, Bool, Num, Str, Arr, Obj } kind = Null;
    double num = 0;
    bool b = false;
    std::string str;
    std::vector<JVal> arr;
    std::vector<std::pair<std::string, JVal>> obj;
    const JVal *get(const char *k) const
    {
        for (auto &kv : obj) if (kv.first == k) return &kv.second;
        return nullptr;
    }
};

struct JParser {
    const char *p, *end;
    bool ok = true;
    void ws() { while (p < end && isspace((unsigned char)*p)) p++; }
    bool lit(const char *s) { size_t n = strlen(s); if ((size_t)(end - p) >= n && !strncmp(p, s, n)) { p += n; return true; } return false; }
    JVal parse()
    {
        JVal v;
        ws();
        if (p >= end) { ok = false; return v; }
        if (*p == '{') {
            v.kind = JVal::Obj; p++; ws();
            if (p < end && *p == '}') { p++; return v; }
            while (ok) {
                ws();
                JVal k = parse();
                if (k.kind != JVal::Str) { ok = false; break; }
                ws();
                if (p >= end || *p != ':') { ok = false; break; }
                p++;
                v.obj.emplace_back(k.str, parse());
                ws();
                if (p < end && *p == ',') { p++; continue; }
                if (p < end && *p == '}') { p++; break; }
                ok = false;
            }
        } else if (*p == '[') {
            v.kind = JVal::Arr; p++; ws();
            if (p < end && *p == ']') { p++; return v; }
            while (ok) {
                v.arr.push_back(parse());
                ws();
                if (p < end && *p == ',') { p++; continue; }
                if (p < end && *p == ']') { p++; break; }
                ok = false;
            }
        } else if (*p == '"') {
            v.kind = JVal::Str; p++;
            while (p < end && *p != '"') {
                if (*p == '\\' && p + 1 < end) { p++; v.str.push_back(*p == 'n' ? '\n' : *p == 't' ? '\t' : *p); }
                else v.str.push_back(*p);
                p++;
            }
            if (p >= end) ok = false; else p++;
        } else if (lit("true")) { v.kind = JVal::Bool; v.b = true; }
        else if (lit("false")) { v.kind = JVal::Bool; v.b = false; }
        else if (lit("null")) { v.kind = JVal::Null; }
        else {
            char *q = nullptr;
            v.kind = JVal::Num; v.num = strtod(p, &q);
            if (q == p) ok = false; else p = q;
        }
        return v;
    }
};

// Both reference parsers read only the first line of the file (getline, Parameter.cpp:29, Src_Rec.cu:32).
bool read_first_line(const std::string &fname, std::string &line)
{
    FILE *fp = fopen(fname.c_str(), "rb");
    if (!fp) return false;
    line.clear();
    int c;
    while ((c = fgetc(fp)) != EOF && c != '\n') line.push_back((char)c);
    fclose(fp);
    return true;
}

// One cached handle per (GPU, grid, mode).  `use` serialises the calls that share a handle (a handle is single-threaded,
// include/sepfwi.h); the shared_ptr keeps it alive for a call that is still running when another call replaces the entry.
struct Cached {
    sepfwi_handle *h = nullptr;
    int maxrec = 0;
    bool dopt_set = false;      // the handle currently carries data-side options ...
    sepfwi_data_options dopt;   // ... these
    std::mutex use;
    ~Cached() { if (h) sepfwi_destroy(h); }
};
std::mutex g_mu;
std::map<std::string, std::shared_ptr<Cached>> g_cache;

bool read_floats(const std::string &fn, float *dst, size_t n)
{
    FILE *fp = fopen(fn.c_str(), "rb");
    if (!fp) return false;
    const size_t got = fread(dst, sizeof(float), n, fp);
    const bool more = got == n && fgetc(fp) != EOF;      // longer than expected: another nrec / nSteps wrote it
    fclose(fp);
    return got == n && !more;
}
bool write_floats(const std::string &fn, const float *src, size_t n)
{
    FILE *fp = fopen(fn.c_str(), "wb");
    if (!fp) return false;
    const size_t put = fwrite(src, sizeof(float), n, fp);
    return fclose(fp) == 0 && put == n;
}

int efail(int code, const std::string &msg);

}  // namespace

extern "C" int sepfwi_set_error_(int code, const char *msg);   // defined in sepfwi.cu

namespace { int efail(int code, const std::string &msg) { return sepfwi_set_error_(code, msg.c_str()); } }

extern "C" int sepfwi_cufd(float *misfit, float *grad_Lambda, float *grad_Mu, float *grad_Den, float *grad_stf,
                           const float *Lambda, const float *Mu, const float *Den, const float *stf,
                           int calc_id, int gpu_id, int group_size, const int *shot_ids, const char *para_fname)
{
    if (calc_id < 0 || calc_id > 2) return efail(SEPFWI_EINVAL, "Invalid calc_id");
    if (!Lambda || !Mu || !Den || !stf || !shot_ids || !para_fname || group_size < 0) return efail(SEPFWI_EINVAL, "null argument");
    std::string line;
    if (!read_first_line(para_fname, line)) return efail(SEPFWI_EIO, std::string("Error opening parameter file ") + para_fname);
    JParser jp{line.data(), line.data() + line.size()};
    JVal para = jp.parse();
    if (!jp.ok || para.kind != JVal::Obj) return efail(SEPFWI_EIO, "parameter file is not a JSON object");
    auto num = [&](const char *k, double &out) { const JVal *v = para.get(k); if (!v || v->kind != JVal::Num) return false; out = v->num; return true; };
    double nz, nx, dz, dx, nSteps, dt, f0, nPml, nPad;
    if (!num("nz", nz) || !num("nx", nx) || !num("dz", dz) || !num("dx", dx) || !num("nSteps", nSteps) || !num("dt", dt) ||
        !num("f0", f0) || !num("nPoints_pml", nPml) || !num("nPad", nPad))
        return efail(SEPFWI_EIO, "parameter file lacks one of nz nx dz dx nSteps dt f0 nPoints_pml nPad");
    const JVal *sv = para.get("survey_fname"), *dd = para.get("data_dir_name");
    if (!sv || sv->kind != JVal::Str || !dd || dd->kind != JVal::Str) return efail(SEPFWI_EIO, "parameter file lacks survey_fname / data_dir_name");
    int fiber = SEPFWI_FIBER_EXX;
    if (const JVal *dc = para.get("das_component")) if (dc->kind == JVal::Str && dc->str == "ezz") fiber = SEPFWI_FIBER_EZZ;
    double max_batch = 0;
    num("max_batch", max_batch);
    int race_compat = 0;
    if (const JVal *rc = para.get("ref_race_compat")) race_compat = (rc->kind == JVal::Bool && rc->b) || (rc->kind == JVal::Num && rc->num != 0);

    if (!read_first_line(sv->str, line)) return efail(SEPFWI_EIO, "Error opening survey file " + sv->str);
    JParser js{line.data(), line.data() + line.size()};
    JVal survey = js.parse();
    if (!js.ok || survey.kind != JVal::Obj) return efail(SEPFWI_EIO, "survey file is not a JSON object");

    const int nS = (int)nSteps, npml = (int)nPml;
    struct ShotData { std::vector<int> zr, xr; std::vector<float> obs, out[4], gstf, w, ws, we, wt; };
    // data-side switches (Parameter.cpp:139-176)
    sepfwi_data_options dopt;
    memset(&dopt, 0, sizeof(dopt));
    if (const JVal *v = para.get("if_win")) dopt.if_win = v->kind == JVal::Bool && v->b;
    if (const JVal *v = para.get("if_src_update")) dopt.if_src_update = v->kind == JVal::Bool && v->b;
    if (const JVal *v = para.get("if_cross_misfit")) dopt.if_cross_misfit = v->kind == JVal::Bool && v->b;
    if (const JVal *v = para.get("filter")) {
        if (v->kind != JVal::Arr || v->arr.size() < 4) return efail(SEPFWI_EIO, "filter must be an array of four frequencies");
        dopt.if_filter = 1;
        for (int k = 0; k < 4; k++) dopt.filter[k] = (float)v->arr[k].num;
    }
    const bool any_dopt = dopt.if_win || dopt.if_filter || dopt.if_cross_misfit || dopt.if_src_update;
    std::vector<ShotData> sd(group_size);
    std::vector<sepfwi_shot> shots(group_size);
    int maxrec = 1;
    for (int i = 0; i < group_size; i++) {
        const std::string key = "shot" + std::to_string(shot_ids[i]);
        const JVal *s = survey.get(key.c_str());
        if (!s || s->kind != JVal::Obj) return efail(SEPFWI_EIO, "survey file lacks " + key);
        const JVal *zs = s->get("z_src"), *xs = s->get("x_src"), *nr = s->get("nrec"), *zr = s->get("z_rec"), *xr = s->get("x_rec");
        if (!zs || !xs || !nr || !zr || !xr || zr->kind != JVal::Arr || xr->kind != JVal::Arr) return efail(SEPFWI_EIO, key + " is incomplete");
        const int nrec = (int)nr->num;
        if ((int)zr->arr.size() < nrec || (int)xr->arr.size() < nrec) return efail(SEPFWI_EIO, key + ": fewer receiver coordinates than nrec");
        sd[i].zr.resize(nrec); sd[i].xr.resize(nrec);
        for (int r = 0; r < nrec; r++) { sd[i].zr[r] = (int)zr->arr[r].num + npml; sd[i].xr[r] = (int)xr->arr[r].num + npml; }   // Src_Rec.cu:108,115
        sepfwi_shot &sh = shots[i];
        memset(&sh, 0, sizeof(sh));
        sh.zs = (int)zs->num + npml; sh.xs = (int)xs->num + npml;                                                               // Src_Rec.cu:87,92
        sh.nrec = nrec; sh.zrec = sd[i].zr.data(); sh.xrec = sd[i].xr.data();
        sh.stf = stf + (size_t)shot_ids[i] * nS;                                                                                // Src_Rec.cu:132
        // extension: per-channel (exx, ezz, exz) sensitivities, "das_sensitivity": [[w0, w1, w2], ...]
        if (const JVal *ds = s->get("das_sensitivity")) {
            if (ds->kind != JVal::Arr || (int)ds->arr.size() < nrec) return efail(SEPFWI_EIO, key + ": das_sensitivity needs nrec rows");
            sd[i].w.resize((size_t)3 * nrec);
            for (int r = 0; r < nrec; r++) {
                const JVal &row = ds->arr[r];
                if (row.kind != JVal::Arr || row.arr.size() < 3) return efail(SEPFWI_EIO, key + ": das_sensitivity rows are [exx, ezz, exz]");
                for (int c = 0; c < 3; c++) sd[i].w[(size_t)3 * r + c] = (float)row.arr[c].num;
            }
            sh.weights = sd[i].w.data();
        }
        auto farr = [&](const char *k, std::vector<float> &dst) -> bool {
            const JVal *a = s->get(k);
            if (!a) return false;
            if (a->kind != JVal::Arr || (int)a->arr.size() < nrec) return false;
            dst.resize(nrec);
            for (int r = 0; r < nrec; r++) dst[r] = (float)a->arr[r].num;
            return true;
        };
        if (dopt.if_win) {                                                                                                     // Src_Rec.cu:145-170
            if (!farr("win_start", sd[i].ws) || !farr("win_end", sd[i].we)) return efail(SEPFWI_EIO, key + ": if_win needs win_start and win_end (nrec entries)");
            sh.win_start = sd[i].ws.data(); sh.win_end = sd[i].we.data();
        }
        if (farr("weights", sd[i].wt)) sh.trace_weights = sd[i].wt.data();                                                     // Src_Rec.cu:176-193
        if (const JVal *sw = s->get("src_weight")) if (sw->kind == JVal::Num) sh.src_weight = (float)sw->num;                // Src_Rec.cu:195-201
        const JVal *rxz = s->get("src_rxz");
        sh.src_rxz = rxz && rxz->kind == JVal::Num ? (float)rxz->num : 1.0f;                                                     // RSXXZZ, utilities.h:21
        maxrec = nrec > maxrec ? nrec : maxrec;
    }

    std::string scratch;
    if (const JVal *sc = para.get("scratch_dir_name")) if (sc->kind == JVal::Str) scratch = sc->str;

    // handle cache
    char keybuf[512];
    snprintf(keybuf, sizeof(keybuf), "%d|%d|%d|%d|%d|%d|%.9g|%.9g|%.9g|%.9g|%d|%d|%d|%d", gpu_id, (int)nz, (int)nx, npml, (int)nPad, nS,
             dz, dx, dt, f0, fiber, calc_id == 1, (int)max_batch, race_compat);
    std::shared_ptr<Cached> ent;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        std::shared_ptr<Cached> &slot = g_cache[keybuf];
        // a cached handle must be able to hold this call's receiver count; a call still using the old one keeps it alive
        if (slot && slot->maxrec < maxrec) slot.reset();
        if (!slot) { slot = std::make_shared<Cached>(); slot->maxrec = maxrec; }
        ent = slot;
    }
    std::lock_guard<std::mutex> use(ent->use);       // held for the whole call: one thread per handle
    if (!ent->h) {
        sepfwi_params p;
        memset(&p, 0, sizeof(p));
        p.nz = (int)nz; p.nx = (int)nx; p.nPml = npml; p.nPad = (int)nPad; p.nSteps = nS;
        p.dz = (float)dz; p.dx = (float)dx; p.dt = (float)dt; p.f0 = (float)f0;
        p.fiber = fiber; p.flavour = SEPFWI_FLAVOUR_CPML; p.max_nrec = ent->maxrec; p.with_adjoint = calc_id == 1; p.ref_race_compat = race_compat;
        int B = (int)max_batch;
        const bool autoB = B <= 0;
        if (autoB) {   // enough concurrent shots to give every launch a few million cells, within memory
            const double cells = (double)(p.nz - p.nPad) * p.nx;
            B = (int)(5.0e7 / cells) + 1;      // measured: throughput still grows from 32 to 64 shots per launch on a 730 k-cell grid
            if (B > 64) B = 64;
            size_t fr = 0, tot = 0;
            if (cudaSetDevice(gpu_id) == cudaSuccess && cudaMemGetInfo(&fr, &tot) == cudaSuccess) {
                const double per = (double)sepfwi_bytes_per_slot(&p);
                while (B > 1 && per * B > 0.6 * (double)fr) B--;
            }
        }
        if (B > group_size && group_size > 0) B = group_size;
        int rc;
        for (;;) {
            p.max_batch = B;
            rc = sepfwi_create(&p, gpu_id, &ent->h);
            if (rc != SEPFWI_ENOMEM || !autoB || B <= 1) break;
            B = B / 2 > 1 ? B / 2 : 1;            // another tenant took the memory between the estimate and the allocation
        }
        if (rc) {
            ent->h = nullptr;
            std::lock_guard<std::mutex> lk(g_mu);
            auto it = g_cache.find(keybuf);
            if (it != g_cache.end() && it->second == ent) g_cache.erase(it);
            return rc;
        }
    }
    sepfwi_handle *h = ent->h;

    int rc = sepfwi_set_model(h, Lambda, Mu, Den, SEPFWI_MEM_HOST, nullptr);
    if (rc) return rc;
    if (calc_id != 2 && (any_dopt != ent->dopt_set || (any_dopt && memcmp(&dopt, &ent->dopt, sizeof(dopt)) != 0))) {
        rc = sepfwi_set_data_options(h, any_dopt ? &dopt : nullptr);      // only when the switches changed: it reallocates scratch
        if (rc) return rc;
        ent->dopt_set = any_dopt; ent->dopt = dopt;
    }

    auto fname = [&](const char *comp, int id) { return dd->str + "/Shot_" + comp + std::to_string(id) + ".bin"; };
    static const char *comps[4] = {"pr", "vx", "vz", "ett"};
    if (calc_id == 2) {
        for (int i = 0; i < group_size; i++)
            for (int c = 0; c < 4; c++) { sd[i].out[c].assign((size_t)shots[i].nrec * nS, 0.f); shots[i].out[c] = sd[i].out[c].data(); }
        rc = sepfwi_forward(h, group_size, shots.data(), SEPFWI_MEM_HOST, nullptr);
        if (rc) return rc;
        for (int i = 0; i < group_size; i++)
            for (int c = 0; c < 4; c++) {
                if (!write_floats(fname(comps[c], shot_ids[i]), sd[i].out[c].data(), sd[i].out[c].size()))
                    return efail(SEPFWI_EIO, "File writing error! " + fname(comps[c], shot_ids[i]));
            }
        return 0;
    }
    for (int i = 0; i < group_size; i++) {
        sd[i].obs.assign((size_t)shots[i].nrec * nS, 0.f);
        const std::string fn = fname("ett", shot_ids[i]);
        if (!read_floats(fn, sd[i].obs.data(), sd[i].obs.size()))
            return efail(SEPFWI_EIO, "File reading error! " + fn + " is missing or does not hold nrec x nSteps = " +
                                         std::to_string(shots[i].nrec) + " x " + std::to_string(nS) + " floats");
        shots[i].obs_ett = sd[i].obs.data();
        if (calc_id == 1 && grad_stf) shots[i].gstf = grad_stf + (size_t)i * nS;   // LOCAL shot index, libCUFD.cu:671-673
    }
    float J = 0.f;
    rc = sepfwi_gradient(h, group_size, shots.data(), calc_id == 1, &J, grad_Lambda, grad_Mu, grad_Den, SEPFWI_MEM_HOST, nullptr);
    if (rc) return rc;
    if (misfit) *misfit = J;
    if (!scratch.empty()) {
        // debug side channel of libCUFD.cu:731-751: pressure residual (obs - syn, sample 0 zeroed as gpuMinus does,
        // utilities.cu:154-167), synthetic pressure and the observed pressure as loaded.  The gradient pass records only the DAS
        // component, so the pressure traces come from one more forward pass -- paid only when the option is set.
        for (int i = 0; i < group_size; i++) {
            for (int c = 0; c < 7; c++) shots[i].out[c] = nullptr;
            sd[i].out[0].assign((size_t)shots[i].nrec * nS, 0.f);
            shots[i].out[0] = sd[i].out[0].data();
        }
        rc = sepfwi_forward(h, group_size, shots.data(), SEPFWI_MEM_HOST, nullptr);
        if (rc) return rc;
        for (int i = 0; i < group_size; i++) {
            const size_t n = (size_t)shots[i].nrec * nS;
            std::vector<float> obs(n), res(n);
            if (!read_floats(fname("pr", shot_ids[i]), obs.data(), n))
                return efail(SEPFWI_EIO, "File reading error! " + fname("pr", shot_ids[i]));
            for (int r = 0; r < shots[i].nrec; r++)
                for (int t = 0; t < nS; t++)
                    res[(size_t)r * nS + t] = t == 0 ? 0.f : obs[(size_t)r * nS + t] - sd[i].out[0][(size_t)r * nS + t];
            const std::string id = std::to_string(shot_ids[i]);
            if (!write_floats(scratch + "/Residual_Shot" + id + ".bin", res.data(), n) ||
                !write_floats(scratch + "/Syn_Shot" + id + ".bin", sd[i].out[0].data(), n) ||
                !write_floats(scratch + "/CondObs_Shot" + id + ".bin", obs.data(), n))
                return efail(SEPFWI_EIO, "File writing error! " + scratch + "/*_Shot" + id + ".bin");
        }
    }
    return 0;
}

// Drop every cached handle (frees device memory); safe to call at any time.
extern "C" int sepfwi_cufd_clear_cache(void)
{
    std::lock_guard<std::mutex> lk(g_mu);
    g_cache.clear();
    return 0;
}
