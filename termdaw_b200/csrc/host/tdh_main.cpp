// termdaw_b200_render <project dir> [--no-normalize] [--device N] [-o out.wav]
// termdaw_b200_render --batch <project dir>... [--no-normalize] [-o name.wav]   (independent projects, one batched call)
// termdaw_b200_render <project dir> --play N [--ahead M] [-o play.f32]   (playback feed: N chunks, raw interleaved f32)
// termdaw_b200_render --dump-project <script.lua>        (seed vectors as JSON; no device needed)
// termdaw_b200_render --dump-midi <file.mid>             (the decoded floww; no device needed)
// termdaw_b200_render --dump-wav <file.wav>              (channels, rate, bits, values, FNV-1a of the decoded f32 stream)
//
// Renders a termdaw project directory: reads project.toml like Config::read (config.rs:10-77), runs the
// Lua-subset reader over the main script, then does what the manual workflow's `refresh` -> `normalize`
// -> `render` commands do (ui_workflow.rs:114-135, state.rs:473-577) and writes the WAV.  Same behaviour
// as `python -m termdaw_b200.cli`; the audio is computed by libtermdaw_b200.so on the GPU — without an
// sm_100 device the context cannot be created and the program says so and exits 1 (no CPU fallback).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <algorithm>
#include <memory>
#include "tdh.h"

int main(int argc, char** argv) {
    using namespace tdh;
    std::string dir, out, dump, dump_midi, dump_wav;
    std::vector<std::string> batch;
    bool normalize = true, batching = false; int device = 0;
    uint64_t play_chunks = 0, ahead = 0;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        if (a == "--no-normalize") normalize = false;
        else if (a == "--batch") batching = true;
        else if (batching && !a.empty() && a[0] != '-') batch.push_back(a);
        else if (a == "--device" && i + 1 < argc) device = atoi(argv[++i]);
        else if ((a == "-o" || a == "--output") && i + 1 < argc) out = argv[++i];
        else if (a == "--play" && i + 1 < argc) play_chunks = strtoull(argv[++i], nullptr, 10);
        else if (a == "--ahead" && i + 1 < argc) ahead = strtoull(argv[++i], nullptr, 10);
        else if (a == "--dump-project" && i + 1 < argc) dump = argv[++i];
        else if (a == "--dump-midi" && i + 1 < argc) dump_midi = argv[++i];
        else if (a == "--dump-wav" && i + 1 < argc) dump_wav = argv[++i];
        else if (!a.empty() && a[0] != '-' && dir.empty()) dir = a;
        else { fprintf(stderr, "usage: %s <project dir> [--no-normalize] [--device N] [-o out.wav] | --dump-project <script.lua>\n", argv[0]); return 2; }
    }
    try {
        if (!dump_wav.empty()) {                               // what SampleBank::add would be handed: header + a digest of the f32 stream
            const DecodedWav w = decode_wav(read_file(dump_wav));
            uint64_t h = 1469598103934665603ull;                // FNV-1a over the f32 bit patterns
            for (float v : w.linear) { uint32_t u; memcpy(&u, &v, 4); for (int k = 0; k < 4; ++k) { h ^= (u >> (8 * k)) & 0xFF; h *= 1099511628211ull; } }
            printf("%u %u %u %zu %llu\n", w.channels, w.sample_rate, w.bits, w.linear.size(), (unsigned long long)h);
            return 0;
        }
        if (!dump_midi.empty()) {                              // the floww a MIDI file decodes to: "t note vel" per event, f32 exact
            const FlowwEvents f = parse_midi(read_file(dump_midi));
            for (size_t i = 0; i < f.t.size(); ++i) printf("%.9g %.9g %.9g\n", (double)f.t[i], (double)f.note[i], (double)f.vel[i]);
            return 0;
        }
        if (!dump.empty()) {
            Project p;
            run_file(dump, p);
            printf("%s\n", p.to_json().c_str());
            return 0;
        }
        if (batching) {
            // the algorithmic-composition batch (BASELINE configs[4]): every directory is one independent project;
            // all of them are loaded, then scanned and rendered in ONE batched call, then written
            char cwd0[4096];
            if (!getcwd(cwd0, sizeof cwd0)) return 1;
            Assets assets;
            std::vector<std::unique_ptr<State>> owned; std::vector<State*> states; std::vector<std::string> dsts;
            for (const std::string& d : batch) {
                if (chdir(cwd0) != 0 || chdir(d.c_str()) != 0) { fprintf(stderr, "TermDaw: could not enter \"%s\".\n", d.c_str()); return 1; }
                const Config cfg = read_config("project.toml");
                Project project;
                run_file(cfg.main, project);
                owned.emplace_back(new State(cfg.project_samplerate, cfg.buffer_length, device, &assets));
                State& st = *owned.back();
                const bool ok = st.refresh(project);
                for (const std::string& m : st.messages) fprintf(stderr, "%s: %s\n", d.c_str(), m.c_str());
                if (!ok) { fprintf(stderr, "TermDaw: project \"%s\" not loaded.\n", d.c_str()); return 1; }
                char here[4096];
                if (!getcwd(here, sizeof here)) return 1;
                const std::string f = out.empty() ? st.output_file : out;
                dsts.push_back(!f.empty() && f[0] == '/' ? f : std::string(here) + "/" + f);
                states.push_back(&st);
            }
            const std::vector<std::vector<uint8_t>> pcms = State::render_batch(states, normalize);
            uint64_t frames = 0;
            for (size_t i = 0; i < states.size(); ++i) { write_wav(dsts[i], pcms[i], states[i]->render_sr, states[i]->bd); frames += states[i]->cs * states[i]->bl; }
            printf("rendered %zu projects, %llu project-rate frames\n", states.size(), (unsigned long long)frames);
            return 0;
        }
        if (dir.empty()) { fprintf(stderr, "usage: %s <project dir> ...\n", argv[0]); return 2; }
        if (chdir(dir.c_str()) != 0) { fprintf(stderr, "TermDaw: could not enter \"%s\".\n", dir.c_str()); return 1; }   // main.rs:39-45
        const Config cfg = read_config("project.toml");
        Project project;
        run_file(cfg.main, project);
        Assets assets;
        State st(cfg.project_samplerate, cfg.buffer_length, device, &assets);
        const bool ok = st.refresh(project);
        for (const std::string& m : st.messages) fprintf(stderr, "%s\n", m.c_str());
        if (!ok) { fprintf(stderr, "TermDaw: project not loaded.\n"); return 1; }
        if (play_chunks > 0) {
            // playback feed to a file: raw interleaved f32 (what the reference queues to SDL), `--play N` chunks in
            // passes of `--ahead M` chunks (default: half a second, the reference's render-ahead)
            if (normalize) st.scan_exact();
            const std::string dst = out.empty() ? "play.f32" : out;
            FILE* f = fopen(dst.c_str(), "wb");
            if (!f) { fprintf(stderr, "TermDaw: could not create \"%s\".\n", dst.c_str()); return 1; }
            const uint64_t step = ahead > 0 ? ahead : (uint64_t)((cfg.project_samplerate / 2 + cfg.buffer_length - 1) / cfg.buffer_length);
            for (uint64_t done = 0; done < play_chunks;) {
                const uint64_t n = std::min<uint64_t>(step, play_chunks - done);
                const std::vector<float> pcm = st.play(n);
                fwrite(pcm.data(), 4, pcm.size(), f);
                done += n;
            }
            fclose(f);
            printf("played %llu chunks -> %s\n", (unsigned long long)play_chunks, dst.c_str());
            return 0;
        }
        if (normalize) st.scan_exact();
        const std::string path = st.render_to_file(out);
        for (const auto& nv : st.normalization_values()) printf("%s: %.9g\n", nv.first.c_str(), (double)nv.second);
        printf("rendered %llu project-rate frames -> %s\n", (unsigned long long)(st.cs * st.bl), path.c_str());
        return 0;
    } catch (const LuaSubsetError& e) {
        fprintf(stderr, "TermDaw: lua error: %s\n", e.what());                  // state.rs:160-164
        return 1;
    } catch (const Error& e) {
        fprintf(stderr, "TermDaw: %s\n", e.what());
        return 1;
    }
}
