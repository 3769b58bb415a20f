// tdh_state.cpp — the host half of `State` (state.rs:27-47) bound to one device context through the C ABI.
// refresh() follows State::refresh (state.rs:50-471) step by step: samples kept when unchanged
// (:201-219), resources (:222-237), flowws always reloaded (:240-250), the graph rebuilt with vertices
// grouped by kind in the FIXED order of state.rs:341-457 — that order defines vertex indices and, through
// connect() order, the f32 summation order — edges in script order (:459), output, check, normalize reset.
#include <math.h>
#include <algorithm>
#include "tdh.h"

namespace tdh {

uint64_t set_length_chunks(uint32_t psr, double seconds, uint32_t bl) {
    // `cs = (psr as f32 * seconds / bl as f32).ceil() as usize` (state.rs:104): f32, one rounding per operation
    volatile float a = (float)psr * (float)seconds;
    volatile float v = a / (float)bl;
    const float c = ceilf(v);
    if (!(c > 0.0f)) return 0;                                 // Rust `as usize` saturates: negative and NaN -> 0
    return (uint64_t)c;
}

State::State(uint32_t project_samplerate, uint32_t buffer_length, int device, Assets* a) : psr(project_samplerate), bl(buffer_length), assets(a) {
    const int rc = TDH_FN(ctx_create)(psr, bl, device, &ctx);
    if (rc < 0 || !ctx) throw Error("ctx_create failed (is a CUDA device visible and the library built?)");
}
State::~State() { if (ctx) TDH_FN(ctx_destroy)(ctx); }

int State::chk(int rc) {
    if (rc < 0) { const char* e = TDH_FN(last_error)(ctx); throw Error(e ? e : "error"); }
    return rc;
}

static std::vector<float> f32s(const std::vector<double>& v) { return std::vector<float>(v.begin(), v.end()); }

bool State::refresh(const Project& p) {
    loaded = false;
    if (p.has_length) cs = set_length_chunks(psr, p.length_seconds, bl);
    if (p.has_render_sr) render_sr = (uint32_t)p.render_samplerate;
    if (p.has_render_bd) bd = (uint32_t)p.render_bitdepth;
    if (p.has_output_file) output_file = p.output_file;
    output_vertex = p.output_vertex;

    // samples: keep what did not change, drop what left the script or changed (state.rs:201-219; SampleBank::mark_dead
    // + refresh, sample.rs:316-336): the device memory of a dropped sample is freed
    const std::vector<SampleSeed>& fresh = p.samples;
    auto has = [](const std::vector<SampleSeed>& v, const SampleSeed& s) { return std::find(v.begin(), v.end(), s) != v.end(); };
    for (const SampleSeed& t : cur_samples)
        if (!has(fresh, t)) {
            say("Info: sample \"" + t.name + "\" will be removed from the sample bank.");
            auto it = sample_index.find(t.name);
            if (it != sample_index.end()) { TDH_FN(sample_remove)(ctx, it->second); sample_index.erase(it); }
        }
    std::vector<std::string> excluded;
    for (const SampleSeed& s : fresh) {
        if (has(cur_samples, s) && sample_index.count(s.name)) continue;
        try {
            const DecodedWav w = assets->wav(s.path);
            if (sample_index.count(s.name)) throw Error("there is already a sample with name \"" + s.name + "\" present.");
            sample_index[s.name] = chk(TDH_FN(sample_add)(ctx, w.linear.data(), w.linear.size(), w.channels, w.sample_rate,
                                                          TDH_FN(sample_mode_from_string)(s.mode.c_str())));
        } catch (const Error& e) { say(e.what()); excluded.push_back(s.name); }
    }
    cur_samples.clear();
    for (const SampleSeed& s : fresh) if (std::find(excluded.begin(), excluded.end(), s.name) == excluded.end()) cur_samples.push_back(s);
    if (!excluded.empty()) return false;                       // do_excluding!, state.rs:188-199

    // resources: diffed as (name, path) tuples like the samples (state.rs:222-237)
    auto has_r = [](const std::vector<NamePath>& v, const NamePath& r) {
        for (const NamePath& q : v) if (q.name == r.name && q.path == r.path) return true;
        return false;
    };
    for (const NamePath& t : cur_resources)
        if (!has_r(p.resources, t)) {
            say("Info: resource \"" + t.name + "\" will be removed from the resource bank.");
            auto it = resource_index.find(t.name);
            if (it != resource_index.end()) { TDH_FN(resource_remove)(ctx, it->second); resource_index.erase(it); }
        }
    std::vector<NamePath> res_loaded;
    for (const NamePath& r : p.resources) {
        if (has_r(cur_resources, r) && resource_index.count(r.name)) { res_loaded.push_back(r); continue; }
        try {
            if (resource_index.count(r.name)) throw Error("there is already a resource with name \"" + r.name + "\" present.");
            const std::vector<uint8_t> b = assets->blob(r.path);
            resource_index[r.name] = chk(TDH_FN(resource_add)(ctx, b.data(), b.size()));
            res_loaded.push_back(r);
        } catch (const Error& e) { say(e.what()); cur_resources = res_loaded; return false; }
    }
    cur_resources = res_loaded;

    TDH_FN(floww_reset)(ctx);                                  // flowws: always reloaded (state.rs:240-250)
    floww_index.clear();
    for (const NamePath& m : p.midis) {
        FlowwEvents f;
        try { f = assets->floww(m.path); } catch (const Error& e) { say(e.what()); return false; }
        try { floww_index[m.name] = chk(TDH_FN(floww_add)(ctx, f.t.data(), f.note.data(), f.vel.data(), f.t.size())); }
        catch (const Error& e) { say(e.what()); return false; }
    }
    for (const std::string& s : p.streams) floww_index[s] = chk(TDH_FN(floww_add)(ctx, nullptr, nullptr, nullptr, 0));

    TDH_FN(graph_reset)(ctx);                                  // graph rebuild in the reference's kind order (state.rs:328-457)
    names.clear(); name_map.clear();
    auto reg = [&](const std::string& name, int idx) { names.push_back(name); name_map[name] = idx; };
    auto need = [&](const std::map<std::string, int>& table, const std::string& key, const std::string& name, const char* category) {
        auto it = table.find(key);
        if (it == table.end()) throw Error(std::string("Could not get ") + category + " index for vertex \"" + name + "\".");
        return it->second;
    };
    try {
        for (const BaseSeed& v : p.sums) reg(v.name, chk(TDH_FN(add_sum)(ctx, (float)v.gain, (float)v.angle)));
        for (const BaseSeed& v : p.norms) reg(v.name, chk(TDH_FN(add_normalize)(ctx, (float)v.gain, (float)v.angle)));
        for (const LoopSeed& v : p.sampleloops)
            reg(v.name, chk(TDH_FN(add_sample_loop)(ctx, (float)v.gain, (float)v.angle, need(sample_index, v.sample, v.name, "sample"))));
        for (const MultiSeed& v : p.samplemultis)
            reg(v.name, chk(TDH_FN(add_sample_multi)(ctx, (float)v.gain, (float)v.angle, need(sample_index, v.sample, v.name, "sample"),
                                                     need(floww_index, v.floww, v.name, "floww"), (int)v.note)));
        for (const LerpSeed& v : p.samplelerps)
            reg(v.name, chk(TDH_FN(add_sample_lerp)(ctx, (float)v.gain, (float)v.angle, need(sample_index, v.sample, v.name, "sample"),
                                                    need(floww_index, v.floww, v.name, "floww"), (int)v.note, (int)v.lerp_len)));
        for (const SineSeed& v : p.debugsines)
            reg(v.name, chk(TDH_FN(add_debug_sine)(ctx, (float)v.gain, (float)v.angle, need(floww_index, v.floww, v.name, "floww"))));
        for (const SynthSeed& v : p.synths) {
            const std::vector<float> a = f32s(v.sq_arr), b = f32s(v.tf_arr), c = f32s(v.tr_arr);
            reg(v.name, chk(TDH_FN(add_synth)(ctx, (float)v.gain, (float)v.angle, need(floww_index, v.floww, v.name, "floww"),
                                              (float)v.sq_vel, (float)v.sq_z, a.data(), (int)a.size(), (float)v.tf_vel, (float)v.tf_z, b.data(), (int)b.size(),
                                              (float)v.tr_vel, c.data(), (int)c.size())));
        }
        for (const SampsynSeed& v : p.sampsyns) {
            const std::vector<float> a = f32s(v.adsr);
            reg(v.name, chk(TDH_FN(add_sampsyn)(ctx, (float)v.gain, (float)v.angle, need(floww_index, v.floww, v.name, "floww"), a.data(), (int)a.size(),
                                                need(resource_index, v.resource, v.name, "resource"))));
        }
        for (const Lv2fxSeed& v : p.lv2fxs)                    // always rejected on this path (TDG_ERR_UNSUPPORTED_VERTEX)
            reg(v.name, chk(TDH_FN(add_lv2fx)(ctx, (float)v.gain, (float)v.angle, (float)v.wet, 0)));
        for (const AdsrSeed& v : p.adsrs) {
            const std::vector<float> a = f32s(v.conf);
            reg(v.name, chk(TDH_FN(add_adsr)(ctx, (float)v.gain, (float)v.angle, (float)v.wet, need(floww_index, v.floww, v.name, "floww"),
                                             v.use_off ? 1 : 0, v.use_max ? 1 : 0, (int)v.note, a.data(), (int)a.size())));
        }
        for (const BandpassSeed& v : p.bandpasses)
            reg(v.name, chk(TDH_FN(add_bandpass)(ctx, (float)v.gain, (float)v.angle, (float)v.wet, (float)v.lo, (float)v.hi, v.pass ? 1 : 0)));
    } catch (const Error& e) { say(e.what()); return false; }

    for (const auto& e : p.edges) {                            // state.rs:459, graph.rs:80-96
        auto a = name_map.find(e.first), b = name_map.find(e.second);
        if (a == name_map.end()) { say("TermDaw: warning: vertex \"" + e.first + "\" cannot be found and thus can't be connected."); continue; }
        if (b == name_map.end()) { say("TermDaw: warning: vertex \"" + e.second + "\" cannot be found and thus can't be connected to."); continue; }
        TDH_FN(connect)(ctx, a->second, b->second);
    }
    auto out = name_map.find(output_vertex);                   // state.rs:461
    if (out != name_map.end()) chk(TDH_FN(set_output)(ctx, out->second));
    std::vector<int> unreachable(4096);
    int n_un = 0;
    const int ok = TDH_FN(check_graph)(ctx, unreachable.data(), (int)unreachable.size(), &n_un);
    for (int i = 0; i < std::min<int>(n_un, (int)unreachable.size()); ++i)
        say("TermDaw: warning: vertex \"" + names[(size_t)unreachable[(size_t)i]] + "\" does not reach output.");
    if (!ok) { say("TermDaw: graph check failed!"); return false; }
    chk(TDH_FN(reset_normalize)(ctx));                         // state.rs:467
    loaded = true;
    return true;
}

void State::scan_exact() { chk(TDH_FN(normalize_scan)(ctx, cs)); }                                  // state.rs:473-475

std::vector<uint8_t> State::render() {                                                               // state.rs:477-577
    const uint64_t cap = TDH_FN(render_bound_bytes)(ctx, cs, render_sr, bd);
    std::vector<uint8_t> out(cap);
    uint64_t frames = 0;
    chk(TDH_FN(render)(ctx, cs, render_sr, bd, out.data(), cap, &frames));
    out.resize(frames * 2 * (bd / 8));
    return out;
}

std::vector<std::vector<uint8_t>> State::render_batch(const std::vector<State*>& states, bool normalize) {
    const int n = (int)states.size();
    std::vector<std::vector<uint8_t>> outs((size_t)n);
    if (n == 0) return outs;
    std::vector<tdg_ctx*> ctxs; std::vector<uint64_t> chunks, caps, frames((size_t)n, 0); std::vector<uint32_t> rsr, bds; std::vector<uint8_t*> ptrs;
    for (int i = 0; i < n; ++i) {
        State* s = states[(size_t)i];
        ctxs.push_back(s->ctx); chunks.push_back(s->cs); rsr.push_back(s->render_sr); bds.push_back(s->bd);
        caps.push_back(TDH_FN(render_bound_bytes)(s->ctx, s->cs, s->render_sr, s->bd));
        outs[(size_t)i].resize(caps.back());
        ptrs.push_back(outs[(size_t)i].data());
    }
    int rc = 0;
#ifndef TDH_ORACLE
    // "normalize" then "render" in ONE pipelined call (scan and render of a project group back to back on a
    // stream, PCM copies of earlier groups under the kernels of later ones)
    if (normalize) rc = tdg_normalize_render_many(ctxs.data(), n, chunks.data(), rsr.data(), bds.data(), ptrs.data(), caps.data(), frames.data());
    else rc = TDH_FN(render_many)(ctxs.data(), n, chunks.data(), rsr.data(), bds.data(), ptrs.data(), caps.data(), frames.data());
#else
    // the oracle (tests only) has the two plain batched calls
    if (normalize) rc = TDH_FN(normalize_scan_many)(ctxs.data(), n, chunks.data());
    if (rc >= 0) rc = TDH_FN(render_many)(ctxs.data(), n, chunks.data(), rsr.data(), bds.data(), ptrs.data(), caps.data(), frames.data());
#endif
    if (rc < 0) {                                              // the failing context holds the message
        for (State* s : states) { const char* e = TDH_FN(last_error)(s->ctx); if (e && *e) throw Error(e); }
        throw Error("batched render failed");
    }
    for (int i = 0; i < n; ++i) outs[(size_t)i].resize(frames[(size_t)i] * 2 * (bds[(size_t)i] / 8));
    return outs;
}

std::vector<float> State::play(uint64_t chunks) {
    std::vector<float> out(chunks * bl * 2);
    uint64_t frames = 0;
    chk(TDH_FN(play)(ctx, chunks, out.data(), chunks * bl, &frames));
    out.resize(frames * 2);
    return out;
}

bool State::feed(const std::string& stream, const FlowwEvents& ev) {
    auto it = floww_index.find(stream);
    if (it == floww_index.end()) { say("stream \"" + stream + "\" is not declared."); return false; }
    chk(TDH_FN(floww_append)(ctx, it->second, ev.t.data(), ev.note.data(), ev.vel.data(), ev.t.size()));
    chk(TDH_FN(floww_set_time)(ctx, TDH_FN(get_time)(ctx)));              // stream_workflow.rs:67-68
    return true;
}

std::string State::render_to_file(const std::string& path) {
    const std::vector<uint8_t> pcm = render();
    const std::string dst = path.empty() ? output_file : path;
    write_wav(dst, pcm, render_sr, bd);
    return dst;
}

std::vector<std::pair<std::string, float>> State::normalization_values() {                          // graph.rs:213-220
    std::vector<std::pair<std::string, float>> out;
    for (size_t i = 0; i < names.size(); ++i) {
        const float v = TDH_FN(get_normalize_value)(ctx, (int)i);
        if (v > 0.0f) out.push_back({names[i], v});
    }
    return out;
}

}  // namespace tdh
