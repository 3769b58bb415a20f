// tdh_assets.cpp — the files either side of the render path: WAV in (hound, sample.rs:231-274), WAV out
// (hound::WavWriter, state.rs:508-531), MIDI -> floww (floww crate, floww.rs:40-48), wavetable blobs
// (bufferbank.rs:26-40), project.toml (config.rs:10-77).  Host I/O only: every number decoded here goes
// to the device untouched (integer PCM is cast `as f32` UNSCALED exactly as sample.rs:266-273 does).
// hound 3.5 and floww 0.1 are not vendored with the reference; the container formats are standard
//  and the MIDI conversion is the documented definition of termdaw_b200/midi.py.
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <algorithm>
#include <tuple>
#include "tdh.h"

namespace tdh {

std::vector<uint8_t> read_file(const std::string& path) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) throw Error("could not open file \"" + path + "\"");
    std::vector<uint8_t> out;
    uint8_t buf[1 << 16];
    size_t n;
    while ((n = fread(buf, 1, sizeof buf, f)) > 0) out.insert(out.end(), buf, buf + n);
    fclose(f);
    return out;
}

static uint16_t rd16(const uint8_t* p) { return (uint16_t)(p[0] | (p[1] << 8)); }
static uint32_t rd32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
static const uint8_t kSubPcm[16] = {0x01, 0x00, 0x00, 0x00, 0x00, 0x00, 0x10, 0x00, 0x80, 0x00, 0x00, 0xAA, 0x00, 0x38, 0x9B, 0x71};

DecodedWav decode_wav(const std::vector<uint8_t>& b) {
    if (b.size() < 12 || memcmp(b.data(), "RIFF", 4) != 0 || memcmp(b.data() + 8, "WAVE", 4) != 0) throw WavError("not a RIFF/WAVE file");
    const uint8_t* fmt = nullptr; size_t fmt_n = 0; const uint8_t* data = nullptr; size_t data_n = 0;
    size_t pos = 12;
    while (pos + 8 <= b.size()) {
        const uint32_t size = rd32(&b[pos + 4]);
        const size_t avail = std::min<size_t>(size, b.size() - (pos + 8));
        if (!fmt && memcmp(&b[pos], "fmt ", 4) == 0) { fmt = &b[pos + 8]; fmt_n = avail; }
        else if (!data && memcmp(&b[pos], "data", 4) == 0) { data = &b[pos + 8]; data_n = avail; }
        pos += 8 + (size_t)size + (size & 1);
    }
    if (!fmt || !data || fmt_n < 16) throw WavError("missing fmt or data chunk");
    uint16_t tag = rd16(fmt);
    const uint16_t channels = rd16(fmt + 2), bits = rd16(fmt + 14);
    const uint32_t rate = rd32(fmt + 4);
    if (tag == 0xFFFE) {                                    // WAVE_FORMAT_EXTENSIBLE
        if (fmt_n < 40) throw WavError("short extensible fmt chunk");
        uint8_t flt[16]; memcpy(flt, kSubPcm, 16); flt[0] = 0x03;
        tag = memcmp(fmt + 24, flt, 16) == 0 ? 3 : 1;
    }
    if (channels == 0) throw WavError("zero channels");
    const size_t nbytes = bits / 8;
    const size_t n = nbytes ? data_n / nbytes : 0;
    DecodedWav w; w.channels = channels; w.sample_rate = rate; w.bits = bits;
    w.linear.resize(n);
    if (tag == 3) {
        if (bits != 32) throw WavError("only 32-bit float WAV is supported");
        for (size_t i = 0; i < n; ++i) { const uint32_t u = rd32(data + 4 * i); memcpy(&w.linear[i], &u, 4); }
    } else if (tag == 1) {
        if (bits == 8) for (size_t i = 0; i < n; ++i) w.linear[i] = (float)((int)data[i] - 128);
        else if (bits == 16) for (size_t i = 0; i < n; ++i) w.linear[i] = (float)(int16_t)rd16(data + 2 * i);
        else if (bits == 24) for (size_t i = 0; i < n; ++i) {
            int32_t v = data[3 * i] | (data[3 * i + 1] << 8) | (data[3 * i + 2] << 16);
            if (v & 0x800000) v -= 1 << 24;
            w.linear[i] = (float)v;
        }
        else if (bits == 32) for (size_t i = 0; i < n; ++i) w.linear[i] = (float)(int32_t)rd32(data + 4 * i);   // i32 `as f32`: round to nearest
        else throw WavError("unsupported bit depth " + std::to_string(bits));
    } else throw WavError("unsupported WAV format tag " + std::to_string(tag));
    return w;
}

static void put16(std::vector<uint8_t>& o, uint32_t v) { o.push_back(v & 0xFF); o.push_back((v >> 8) & 0xFF); }
static void put32(std::vector<uint8_t>& o, uint32_t v) { put16(o, v & 0xFFFF); put16(o, v >> 16); }
static void puts4(std::vector<uint8_t>& o, const char* s) { o.insert(o.end(), s, s + 4); }

// hound::WavWriter: 16-byte PCM fmt chunk up to 16 bit, WAVE_FORMAT_EXTENSIBLE (40 bytes) above
std::vector<uint8_t> wav_header(uint64_t data_bytes, uint32_t sample_rate, uint32_t bitdepth, uint32_t channels) {
    const uint32_t block = channels * (bitdepth / 8);
    std::vector<uint8_t> fmt;
    if (bitdepth <= 16) {
        put16(fmt, 1); put16(fmt, channels); put32(fmt, sample_rate); put32(fmt, sample_rate * block); put16(fmt, block); put16(fmt, bitdepth);
    } else {
        const uint32_t mask = channels == 1 ? 0x4 : channels == 2 ? 0x3 : 0;
        put16(fmt, 0xFFFE); put16(fmt, channels); put32(fmt, sample_rate); put32(fmt, sample_rate * block); put16(fmt, block); put16(fmt, bitdepth);
        put16(fmt, 22); put16(fmt, bitdepth); put32(fmt, mask);
        fmt.insert(fmt.end(), kSubPcm, kSubPcm + 16);
    }
    std::vector<uint8_t> o;
    puts4(o, "RIFF"); put32(o, (uint32_t)(4 + (8 + fmt.size()) + (8 + data_bytes))); puts4(o, "WAVE");
    puts4(o, "fmt "); put32(o, (uint32_t)fmt.size()); o.insert(o.end(), fmt.begin(), fmt.end());
    puts4(o, "data"); put32(o, (uint32_t)data_bytes);
    return o;
}

void write_wav(const std::string& path, const std::vector<uint8_t>& pcm, uint32_t sample_rate, uint32_t bitdepth) {
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) throw Error("could not create file \"" + path + "\"");
    const std::vector<uint8_t> h = wav_header(pcm.size(), sample_rate, bitdepth, 2);
    const bool ok = fwrite(h.data(), 1, h.size(), f) == h.size() && (pcm.empty() || fwrite(pcm.data(), 1, pcm.size(), f) == pcm.size());
    if (fclose(f) != 0 || !ok) throw Error("could not write file \"" + path + "\"");
}

// ---- Standard MIDI File -> floww events (definition: termdaw_b200/midi.py) ---------------------------
namespace {
struct NoteEv { uint64_t tick; int track, seq; int key, vel; };
struct TempoEv { uint64_t tick; int track, seq; uint32_t us; };

uint32_t vlq(const uint8_t* b, size_t n, size_t& i) {
    uint32_t v = 0;
    for (int k = 0; k < 4; ++k) {
        if (i >= n) throw MidiError("truncated variable-length quantity");
        const uint8_t c = b[i++];
        v = (v << 7) | (c & 0x7F);
        if (!(c & 0x80)) return v;
    }
    throw MidiError("variable-length quantity longer than 4 bytes");
}

uint64_t parse_track(const uint8_t* b, size_t n, int track, std::vector<NoteEv>& notes, std::vector<TempoEv>& tempos) {
    size_t i = 0; uint64_t tick = 0; uint8_t status = 0; int seq = 0;
    while (i < n) {
        tick += vlq(b, n, i);
        if (i >= n) throw MidiError("track ends inside an event");
        const uint8_t c = b[i];
        if (c == 0xFF) {
            if (i + 1 >= n) throw MidiError("truncated meta event");
            const uint8_t kind = b[i + 1];
            size_t j = i + 2;
            const uint32_t ln = vlq(b, n, j);
            if (j + ln > n) throw MidiError("truncated meta event");
            i = j + ln; status = 0;
            if (kind == 0x51 && ln == 3) tempos.push_back({tick, track, seq, ((uint32_t)b[j] << 16) | ((uint32_t)b[j + 1] << 8) | b[j + 2]});
            else if (kind == 0x2F) break;
        } else if (c == 0xF0 || c == 0xF7) {
            size_t j = i + 1;
            const uint32_t ln = vlq(b, n, j);
            if (j + ln > n) throw MidiError("truncated sysex event");
            i = j + ln; status = 0;
        } else {
            if (c & 0x80) { status = c; ++i; }
            else if (!status) throw MidiError("data byte without a running status");
            const uint8_t hi = status & 0xF0;
            if (hi == 0x80 || hi == 0x90 || hi == 0xA0 || hi == 0xB0 || hi == 0xE0) {
                if (i + 2 > n) throw MidiError("truncated channel event");
                const int d0 = b[i], d1 = b[i + 1];
                i += 2;
                if (hi == 0x90 && d1 > 0) notes.push_back({tick, track, seq, d0 & 0x7F, d1 & 0x7F});
                else if (hi == 0x80 || hi == 0x90) notes.push_back({tick, track, seq, d0 & 0x7F, 0});
            } else if (hi == 0xC0 || hi == 0xD0) {
                if (i + 1 > n) throw MidiError("truncated channel event");
                i += 1;
            } else if (status == 0xF1 || status == 0xF3) i += 1;
            else if (status == 0xF2) i += 2;
        }
        ++seq;
    }
    return tick;
}
}  // namespace

FlowwEvents parse_midi(const std::vector<uint8_t>& d) {
    if (d.size() < 14 || memcmp(d.data(), "MThd", 4) != 0) throw MidiError("not a Standard MIDI File (no MThd)");
    auto be16 = [&](size_t p) { return (uint32_t)((d[p] << 8) | d[p + 1]); };
    auto be32 = [&](size_t p) { return ((uint32_t)d[p] << 24) | ((uint32_t)d[p + 1] << 16) | ((uint32_t)d[p + 2] << 8) | d[p + 3]; };
    const uint32_t hlen = be32(4), fmt = be16(8), ntrks = be16(10), division = be16(12);
    if (hlen < 6) throw MidiError("bad MThd length");
    size_t i = 8 + (size_t)hlen;
    std::vector<NoteEv> notes; std::vector<TempoEv> tempos;
    uint64_t offset = 0; uint32_t k = 0;
    while (i + 8 <= d.size() && k < ntrks) {
        const uint32_t ln = be32(i + 4);
        if (i + 8 + (size_t)ln > d.size()) throw MidiError("truncated chunk");
        if (memcmp(&d[i], "MTrk", 4) == 0) {
            const size_t n0 = notes.size(), t0 = tempos.size();
            const uint64_t end = parse_track(&d[i + 8], ln, (int)k, notes, tempos);
            if (fmt == 2) {
                for (size_t q = n0; q < notes.size(); ++q) notes[q].tick += offset;
                for (size_t q = t0; q < tempos.size(); ++q) tempos[q].tick += offset;
                offset += end;
            }
            ++k;
        }
        i += 8 + (size_t)ln;
    }
    if (k == 0) throw MidiError("no MTrk chunk");
    auto key = [](const auto& e) { return std::make_tuple(e.tick, e.track, e.seq); };
    std::stable_sort(notes.begin(), notes.end(), [&](const NoteEv& a, const NoteEv& b) { return key(a) < key(b); });
    std::stable_sort(tempos.begin(), tempos.end(), [&](const TempoEv& a, const TempoEv& b) { return key(a) < key(b); });
    std::vector<double> secs(notes.size());
    if (division & 0x8000) {                                  // SMPTE
        const int fps = 256 - (int)(division >> 8);
        const double fps_f = fps == 29 ? 29.97 : (double)fps;
        const double per_tick = 1.0 / (fps_f * (double)(division & 0xFF));
        for (size_t q = 0; q < notes.size(); ++q) secs[q] = (double)notes[q].tick * per_tick;
    } else {
        if (division == 0) throw MidiError("division of 0 ticks per quarter note");
        std::vector<uint64_t> seg_tick{0}; std::vector<double> seg_sec{0.0}; std::vector<uint32_t> seg_us{500000};
        for (const TempoEv& t : tempos) {
            if (t.tick == seg_tick.back()) { seg_us.back() = t.us; continue; }
            seg_sec.push_back(seg_sec.back() + (double)(t.tick - seg_tick.back()) * (double)seg_us.back() / (double)division / 1e6);
            seg_tick.push_back(t.tick); seg_us.push_back(t.us);
        }
        size_t s = 0;
        for (size_t q = 0; q < notes.size(); ++q) {
            while (s + 1 < seg_tick.size() && seg_tick[s + 1] <= notes[q].tick) ++s;
            secs[q] = seg_sec[s] + (double)(notes[q].tick - seg_tick[s]) * (double)seg_us[s] / (double)division / 1e6;
        }
    }
    FlowwEvents f;
    for (size_t q = 0; q < notes.size(); ++q) {
        f.t.push_back((float)secs[q]);
        f.note.push_back((float)notes[q].key);
        f.vel.push_back((float)notes[q].vel / 127.0f);
    }
    return f;
}

// ---- disk assets -------------------------------------------------------------------------------------
DecodedWav Assets::wav(const std::string& path) {
    std::vector<uint8_t> b;
    try { b = read_file(path); } catch (const Error&) { throw Error("TermDaw: SampleBank: could not open file \"" + path + "\"."); }   // sample.rs:231-240
    return decode_wav(b);
}
FlowwEvents Assets::floww(const std::string& path) {
    // a decoded event stream next to the file ("<file>.floww.f32": little-endian f32 triples t, note, vel)
    // takes precedence: it lets synthetic flowws reach the native host without a MIDI encoder in between
    try {
        const std::vector<uint8_t> side = read_file(path + ".floww.f32");
        FlowwEvents f;
        for (size_t i = 0; i + 12 <= side.size(); i += 12) {
            float v[3]; memcpy(v, &side[i], 12);
            f.t.push_back(v[0]); f.note.push_back(v[1]); f.vel.push_back(v[2]);
        }
        return f;
    } catch (const Error&) {}
    try { return parse_midi(read_file(path)); }
    catch (const Error&) { throw Error("Could not read midi file: \"" + path + "\"."); }             // floww.rs:45-46
}
std::vector<uint8_t> Assets::blob(const std::string& path) { return read_file(path); }

// ---- project.toml: the three settings Config::read uses (config.rs:10-77) ------------------------------
Config read_config(const std::string& path) {
    const std::vector<uint8_t> b = read_file(path);
    Config c;
    std::string section, line;
    auto trim = [](std::string s) {
        const size_t a = s.find_first_not_of(" \t\r"), z = s.find_last_not_of(" \t\r");
        return a == std::string::npos ? std::string() : s.substr(a, z - a + 1);
    };
    auto handle = [&](std::string ln) {
        bool in_str = false;                                   // strip a trailing comment (outside strings)
        for (size_t i = 0; i < ln.size(); ++i) { if (ln[i] == '"') in_str = !in_str; else if (ln[i] == '#' && !in_str) { ln.resize(i); break; } }
        ln = trim(ln);
        if (ln.empty()) return;
        if (ln.front() == '[' && ln.back() == ']') { section = trim(ln.substr(1, ln.size() - 2)); return; }
        const size_t eq = ln.find('=');
        if (eq == std::string::npos || section != "settings") return;
        const std::string key = trim(ln.substr(0, eq));
        std::string val = trim(ln.substr(eq + 1));
        if (val.size() >= 2 && val.front() == '"' && val.back() == '"') val = val.substr(1, val.size() - 2);
        if (key == "main") c.main = val;
        else if (key == "workflow") c.workflow = val;
        else if (key == "buffer_length") c.buffer_length = (uint32_t)strtoul(val.c_str(), nullptr, 10);
        else if (key == "project_samplerate") c.project_samplerate = (uint32_t)strtoul(val.c_str(), nullptr, 10);
    };
    for (uint8_t ch : b) { if (ch == '\n') { handle(line); line.clear(); } else line += (char)ch; }
    handle(line);
    if (c.main.empty()) throw Error("project.toml: [settings] main is missing");
    return c;
}

}  // namespace tdh
