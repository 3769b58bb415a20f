// tdh_lua.cpp — the subset of Lua 5.3 that termdaw project scripts use (state.rs:83-164 runs them in mlua).
//
// Every registered global only appends a tuple to a seed vector (state.rs:103-157), so a project script
// is a list of calls with literal arguments plus a few variables (project.lua, examples/*.lua):
//
//   statements   name = expr | local name = expr | f(args) | for i = a, b [, step] do ... end
//   expressions  numbers, "strings" / 'strings', true / false / nil, { e, e, ... } (array tables), names,
//                unary -, binary + - * / % .. with Lua precedence, parentheses
//   comments     -- line   and   --[[ block ]]          optional ';' after a statement
//
// Anything else throws LuaSubsetError (the reference prints the mlua error and leaves the project
// unloaded, state.rs:160-164).  Same grammar as termdaw_b200/lua_subset.py; tests/test_host_native.py
// checks that both produce the same seed vectors from the reference's own scripts.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <functional>
#include "tdh.h"

namespace tdh {
namespace {

struct Value {
    enum Kind { NIL, BOOL, INT, FLOAT, STR, LIST } kind = NIL;
    bool b = false; int64_t i = 0; double f = 0; std::string s; std::vector<Value> list;
    static Value Int(int64_t v) { Value x; x.kind = INT; x.i = v; return x; }
    static Value Float(double v) { Value x; x.kind = FLOAT; x.f = v; return x; }
    static Value Str(std::string v) { Value x; x.kind = STR; x.s = std::move(v); return x; }
    static Value Bool(bool v) { Value x; x.kind = BOOL; x.b = v; return x; }
    bool is_num() const { return kind == INT || kind == FLOAT; }
    double num() const { return kind == INT ? (double)i : f; }
};

struct Token { enum Kind { NUM, STR, NAME, KW, OP, END } kind; Value num; std::string text; int line; };

[[noreturn]] void fail(int line, const std::string& m) { throw LuaSubsetError("line " + std::to_string(line) + ": " + m); }

bool is_keyword(const std::string& s) {
    static const char* kw[] = {"local", "for", "do", "end", "true", "false", "nil"};
    for (const char* k : kw) if (s == k) return true;
    return false;
}

std::vector<Token> tokenize(const std::string& t) {
    std::vector<Token> out;
    size_t p = 0; int line = 1;
    const size_t n = t.size();
    auto isname0 = [](char c) { return isalpha((unsigned char)c) || c == '_'; };
    auto isname = [](char c) { return isalnum((unsigned char)c) || c == '_'; };
    while (p < n) {
        const char c = t[p];
        if (c == '\n') { ++line; ++p; continue; }
        if (isspace((unsigned char)c)) { ++p; continue; }
        if (c == '-' && p + 1 < n && t[p + 1] == '-') {
            if (t.compare(p, 4, "--[[") == 0) {
                const size_t e = t.find("]]", p + 4);
                if (e != std::string::npos) {
                    for (size_t q = p; q < e; ++q) if (t[q] == '\n') ++line;
                    p = e + 2; continue;
                }
            }
            while (p < n && t[p] != '\n') ++p;
            continue;
        }
        if (isdigit((unsigned char)c) || (c == '.' && p + 1 < n && isdigit((unsigned char)t[p + 1]))) {
            size_t q = p;
            Token tok{Token::NUM, Value(), "", line};
            if (c == '0' && q + 1 < n && (t[q + 1] == 'x' || t[q + 1] == 'X') && q + 2 < n && isxdigit((unsigned char)t[q + 2])) {
                q += 2;
                while (q < n && isxdigit((unsigned char)t[q])) ++q;
                tok.num = Value::Int((int64_t)strtoull(t.substr(p + 2, q - p - 2).c_str(), nullptr, 16));
            } else {
                bool is_int = true;
                while (q < n && isdigit((unsigned char)t[q])) ++q;
                if (q < n && t[q] == '.') { is_int = false; ++q; while (q < n && isdigit((unsigned char)t[q])) ++q; }
                if (q < n && (t[q] == 'e' || t[q] == 'E')) {
                    size_t r = q + 1;
                    if (r < n && (t[r] == '+' || t[r] == '-')) ++r;
                    if (r < n && isdigit((unsigned char)t[r])) { is_int = false; q = r; while (q < n && isdigit((unsigned char)t[q])) ++q; }
                }
                const std::string txt = t.substr(p, q - p);
                tok.num = is_int ? Value::Int((int64_t)strtoll(txt.c_str(), nullptr, 10)) : Value::Float(strtod(txt.c_str(), nullptr));
            }
            tok.text = t.substr(p, q - p);
            out.push_back(tok); p = q; continue;
        }
        if (c == '"' || c == '\'') {
            size_t q = p + 1; std::string body;
            while (q < n && t[q] != c) {
                if (t[q] == '\\' && q + 1 < n) {
                    const char e = t[q + 1];
                    body += e == 'n' ? '\n' : e == 't' ? '\t' : e == 'r' ? '\r' : e == '0' ? '\0' : e;
                    q += 2;
                } else { if (t[q] == '\n') ++line; body += t[q++]; }
            }
            if (q >= n) fail(line, "unfinished string");
            out.push_back({Token::STR, Value(), body, line}); p = q + 1; continue;
        }
        if (isname0(c)) {
            size_t q = p;
            while (q < n && isname(t[q])) ++q;
            const std::string w = t.substr(p, q - p);
            out.push_back({is_keyword(w) ? Token::KW : Token::NAME, Value(), w, line}); p = q; continue;
        }
        if (c == '.' && p + 1 < n && t[p + 1] == '.') { out.push_back({Token::OP, Value(), "..", line}); p += 2; continue; }
        if (strchr("=(){},;+-*/%", c)) { out.push_back({Token::OP, Value(), std::string(1, c), line}); ++p; continue; }
        fail(line, std::string("unexpected character '") + c + "'");
    }
    out.push_back({Token::END, Value(), "", line});
    return out;
}

std::string tostr(const Value& x) {
    char buf[64];
    if (x.kind == Value::STR) return x.s;
    if (x.kind == Value::INT) { snprintf(buf, sizeof buf, "%lld", (long long)x.i); return buf; }
    if (x.kind == Value::FLOAT) {
        if (x.f == floor(x.f) && fabs(x.f) < 1e15) snprintf(buf, sizeof buf, "%.1f", x.f);
        else snprintf(buf, sizeof buf, "%.14g", x.f);                   // Lua's LUAI_NUMFFORMAT
        return buf;
    }
    throw LuaSubsetError("attempt to concatenate a non-string, non-number value");
}

Value arith(const std::string& op, const Value& a, const Value& b, int line) {
    if (!a.is_num() || !b.is_num()) fail(line, "attempt to perform arithmetic on a non-number");
    const bool ints = a.kind == Value::INT && b.kind == Value::INT;
    if (op == "+") return ints ? Value::Int(a.i + b.i) : Value::Float(a.num() + b.num());
    if (op == "-") return ints ? Value::Int(a.i - b.i) : Value::Float(a.num() - b.num());
    if (op == "*") return ints ? Value::Int(a.i * b.i) : Value::Float(a.num() * b.num());
    if (op == "/") return Value::Float(a.num() / b.num());              // Lua 5.3: '/' is always float division
    if (ints) {
        if (b.i == 0) fail(line, "attempt to perform 'n%%0'");
        int64_t m = a.i % b.i;
        if (m != 0 && ((m ^ b.i) < 0)) m += b.i;                         // floored modulo
        return Value::Int(m);
    }
    if (b.num() == 0) fail(line, "attempt to perform 'n%%0'");
    return Value::Float(a.num() - floor(a.num() / b.num()) * b.num());
}

// ---- argument conversion (what mlua does for the closure signatures of state.rs:103-157) -------
struct Args {
    const std::vector<Value>& v; const std::string& fn; int line;
    void arity(size_t lo, size_t hi) const {
        if (v.size() < lo || v.size() > hi) fail(line, "bad arguments to '" + fn + "': expected " + std::to_string(lo) +
                                                 (hi != lo ? ".." + std::to_string(hi) : "") + " arguments, got " + std::to_string(v.size()));
    }
    [[noreturn]] void bad(size_t i, const char* want) const { fail(line, "bad arguments to '" + fn + "': argument " + std::to_string(i + 1) + " is not " + want); }
    std::string str(size_t i) const {
        if (v[i].kind == Value::STR) return v[i].s;
        if (v[i].is_num()) return tostr(v[i]);
        bad(i, "a string");
    }
    double num(size_t i) const { if (!v[i].is_num()) bad(i, "a number"); return v[i].num(); }
    int64_t integer(size_t i) const {
        if (v[i].kind == Value::INT) return v[i].i;
        if (v[i].kind == Value::FLOAT) return (int64_t)v[i].f;          // truncates, like the Python mirror's int()
        bad(i, "an integer");
    }
    bool boolean(size_t i) const { return !(v[i].kind == Value::NIL || (v[i].kind == Value::BOOL && !v[i].b)); }   // Lua truthiness
    std::vector<double> floats(size_t i) const {
        if (v[i].kind != Value::LIST) bad(i, "a table");
        std::vector<double> out;
        for (const Value& e : v[i].list) { if (!e.is_num()) bad(i, "a table of numbers"); out.push_back(e.num()); }
        return out;
    }
};

using Fn = std::function<void(const Args&)>;

std::map<std::string, Fn> globals_of(Project& p) {
    std::map<std::string, Fn> g;
    g["set_length"] = [&p](const Args& a) { a.arity(1, 1); p.set_length(a.num(0)); };
    g["set_render_samplerate"] = [&p](const Args& a) { a.arity(1, 1); p.set_render_samplerate(a.integer(0)); };
    g["set_render_bitdepth"] = [&p](const Args& a) { a.arity(1, 1); p.set_render_bitdepth(a.integer(0)); };
    g["set_output_file"] = [&p](const Args& a) { a.arity(1, 1); p.set_output_file(a.str(0)); };
    g["load_sample"] = [&p](const Args& a) { a.arity(2, 3); p.load_sample(a.str(0), a.str(1), a.v.size() > 2 ? a.str(2) : ""); };
    g["load_resource"] = [&p](const Args& a) { a.arity(2, 2); p.load_resource(a.str(0), a.str(1)); };
    g["load_midi_floww"] = [&p](const Args& a) { a.arity(2, 2); p.load_midi_floww(a.str(0), a.str(1)); };
    g["declare_stream"] = [&p](const Args& a) { a.arity(1, 1); p.declare_stream(a.str(0)); };
    g["load_lv2"] = [&p](const Args& a) { a.arity(2, 2); p.load_lv2(a.str(0), a.str(1)); };
    g["parameter"] = [&p](const Args& a) { a.arity(3, 3); p.parameter(a.str(0), a.str(1), a.num(2)); };
    g["add_sum"] = [&p](const Args& a) { a.arity(3, 3); p.add_sum(a.str(0), a.num(1), a.num(2)); };
    g["add_normalize"] = [&p](const Args& a) { a.arity(3, 3); p.add_normalize(a.str(0), a.num(1), a.num(2)); };
    g["add_sampleloop"] = [&p](const Args& a) { a.arity(4, 4); p.add_sampleloop(a.str(0), a.num(1), a.num(2), a.str(3)); };
    g["add_sample_multi"] = [&p](const Args& a) { a.arity(6, 6); p.add_sample_multi(a.str(0), a.num(1), a.num(2), a.str(3), a.str(4), a.integer(5)); };
    g["add_sample_lerp"] = [&p](const Args& a) { a.arity(7, 7); p.add_sample_lerp(a.str(0), a.num(1), a.num(2), a.str(3), a.str(4), a.integer(5), a.integer(6)); };
    g["add_debug_sine"] = [&p](const Args& a) { a.arity(4, 4); p.add_debug_sine(a.str(0), a.num(1), a.num(2), a.str(3)); };
    g["add_synth"] = [&p](const Args& a) { a.arity(12, 12); p.add_synth(a.str(0), a.num(1), a.num(2), a.str(3), a.num(4), a.num(5), a.floats(6),
                                                                        a.num(7), a.num(8), a.floats(9), a.num(10), a.floats(11)); };
    g["add_sampsyn"] = [&p](const Args& a) { a.arity(6, 6); p.add_sampsyn(a.str(0), a.num(1), a.num(2), a.str(3), a.floats(4), a.str(5)); };
    g["add_lv2fx"] = [&p](const Args& a) { a.arity(5, 5); p.add_lv2fx(a.str(0), a.num(1), a.num(2), a.num(3), a.str(4)); };
    g["add_adsr"] = [&p](const Args& a) { a.arity(9, 9); p.add_adsr(a.str(0), a.num(1), a.num(2), a.num(3), a.str(4), a.boolean(5), a.boolean(6),
                                                                    a.integer(7), a.floats(8)); };
    g["add_bandpass"] = [&p](const Args& a) { a.arity(7, 7); p.add_bandpass(a.str(0), a.num(1), a.num(2), a.num(3), a.num(4), a.num(5), a.boolean(6)); };
    g["connect"] = [&p](const Args& a) { a.arity(2, 2); p.connect(a.str(0), a.str(1)); };
    g["set_output"] = [&p](const Args& a) { a.arity(1, 1); p.set_output(a.str(0)); };
    // north-star paraphrases (BASELINE.json) kept as aliases of the real names
    g["add_samplemulti"] = g["add_sample_multi"]; g["add_samplelerp"] = g["add_sample_lerp"];
    g["add_wavetable"] = g["add_sampsyn"]; g["add_envelope"] = g["add_adsr"];
    return g;
}

struct Interp {
    std::vector<Token> t; size_t i = 0;
    std::map<std::string, Fn> functions;
    std::map<std::string, Value> globals;
    std::vector<std::map<std::string, Value>> scopes;

    const Token& peek() const { return t[i]; }
    const Token& next() { return t[i++]; }
    bool accept(Token::Kind k, const char* val = nullptr) {
        if (peek().kind == k && (!val || peek().text == val)) { ++i; return true; }
        return false;
    }
    const Token& expect(Token::Kind k, const char* val = nullptr) {
        if (peek().kind != k || (val && peek().text != val))
            fail(peek().line, std::string("expected ") + (val ? val : k == Token::NAME ? "name" : "token") + ", found '" + peek().text + "'");
        return next();
    }
    Value lookup(const std::string& name, int line) {
        for (size_t s = scopes.size(); s-- > 0;) { auto it = scopes[s].find(name); if (it != scopes[s].end()) return it->second; }
        auto it = globals.find(name);
        if (it != globals.end()) return it->second;
        if (functions.count(name)) fail(line, "function '" + name + "' used as a value (outside the supported subset)");
        return Value();                                      // undefined global is nil in Lua
    }
    // concat < additive < multiplicative < unary < primary
    Value expr() {
        Value left = additive();
        if (accept(Token::OP, "..")) { Value right = expr(); return Value::Str(tostr(left) + tostr(right)); }   // right associative
        return left;
    }
    Value additive() {
        Value v = multiplicative();
        for (;;) {
            const Token& k = peek();
            if (k.kind == Token::OP && (k.text == "+" || k.text == "-")) { const std::string op = k.text; const int line = k.line; next(); v = arith(op, v, multiplicative(), line); }
            else return v;
        }
    }
    Value multiplicative() {
        Value v = unary();
        for (;;) {
            const Token& k = peek();
            if (k.kind == Token::OP && (k.text == "*" || k.text == "/" || k.text == "%")) { const std::string op = k.text; const int line = k.line; next(); v = arith(op, v, unary(), line); }
            else return v;
        }
    }
    Value unary() {
        if (peek().kind == Token::OP && peek().text == "-") {
            const int line = next().line;
            Value x = unary();
            if (!x.is_num()) fail(line, "unary minus on a non-number");
            return x.kind == Value::INT ? Value::Int(-x.i) : Value::Float(-x.f);
        }
        return primary();
    }
    Value primary() {
        const Token k = next();
        if (k.kind == Token::NUM) return k.num;
        if (k.kind == Token::STR) return Value::Str(k.text);
        if (k.kind == Token::KW && (k.text == "true" || k.text == "false")) return Value::Bool(k.text == "true");
        if (k.kind == Token::KW && k.text == "nil") return Value();
        if (k.kind == Token::NAME) {
            if (peek().kind == Token::OP && peek().text == "(") fail(k.line, "call of '" + k.text + "' inside an expression is outside the supported subset");
            return lookup(k.text, k.line);
        }
        if (k.kind == Token::OP && k.text == "(") { Value x = expr(); expect(Token::OP, ")"); return x; }
        if (k.kind == Token::OP && k.text == "{") {
            Value l; l.kind = Value::LIST;
            while (!accept(Token::OP, "}")) {
                l.list.push_back(expr());
                if (!(accept(Token::OP, ",") || accept(Token::OP, ";"))) { expect(Token::OP, "}"); break; }
            }
            return l;
        }
        fail(k.line, "unexpected token '" + k.text + "'");
    }
    void block(bool until_end) {
        for (;;) {
            const Token& k = peek();
            if (k.kind == Token::END) { if (until_end) fail(k.line, "'end' expected"); return; }
            if (k.kind == Token::KW && k.text == "end") { if (!until_end) fail(k.line, "unexpected 'end'"); next(); return; }
            statement();
        }
    }
    void skip_block() {
        int depth = 1;
        while (depth) {
            const Token& k = next();
            if (k.kind == Token::END) fail(k.line, "'end' expected");
            if (k.kind == Token::KW && k.text == "do") ++depth;
            else if (k.kind == Token::KW && k.text == "end") --depth;
        }
    }
    void assign(const std::string& name, const Value& v) {
        for (size_t s = scopes.size(); s-- > 0;) { auto it = scopes[s].find(name); if (it != scopes[s].end()) { it->second = v; return; } }
        globals[name] = v;
    }
    void statement() {
        const Token k = next();
        if (k.kind == Token::OP && k.text == ";") return;
        if (k.kind == Token::KW && k.text == "local") {
            const std::string name = expect(Token::NAME).text;
            Value val;
            if (accept(Token::OP, "=")) val = expr();
            (scopes.empty() ? globals : scopes.back())[name] = val;
            accept(Token::OP, ";");
            return;
        }
        if (k.kind == Token::KW && k.text == "for") {
            const std::string var = expect(Token::NAME).text;
            expect(Token::OP, "=");
            Value a = expr(); expect(Token::OP, ","); Value b = expr();
            Value step = Value::Int(1);
            if (accept(Token::OP, ",")) step = expr();
            expect(Token::KW, "do");
            if (!a.is_num() || !b.is_num() || !step.is_num()) fail(k.line, "'for' bounds must be numbers");
            if (step.num() == 0) fail(k.line, "'for' step is zero");
            const size_t body = i;
            bool ran = false;
            Value x = a;
            while ((step.num() > 0 && x.num() <= b.num()) || (step.num() < 0 && x.num() >= b.num())) {
                i = body;
                scopes.push_back({{var, x}});
                block(true);
                scopes.pop_back();
                ran = true;
                x = arith("+", x, step, k.line);
            }
            if (!ran) { i = body; skip_block(); }
            accept(Token::OP, ";");
            return;
        }
        if (k.kind == Token::NAME) {
            if (accept(Token::OP, "(")) {
                std::vector<Value> args;
                if (!accept(Token::OP, ")")) {
                    for (;;) { args.push_back(expr()); if (accept(Token::OP, ")")) break; expect(Token::OP, ","); }
                }
                auto it = functions.find(k.text);
                if (it == functions.end()) fail(k.line, "attempt to call a nil value (global '" + k.text + "')");
                it->second(Args{args, k.text, k.line});
                accept(Token::OP, ";");
                return;
            }
            if (accept(Token::OP, "=")) { assign(k.text, expr()); accept(Token::OP, ";"); return; }
        }
        fail(k.line, "unsupported statement starting with '" + k.text + "'");
    }
};

}  // namespace

void run_script(const std::string& text, Project& p) {
    Interp it;
    it.t = tokenize(text);
    it.functions = globals_of(p);
    it.block(false);
}

void run_file(const std::string& path, Project& p) {
    const std::vector<uint8_t> b = read_file(path);
    run_script(std::string(b.begin(), b.end()), p);
}

// ---- JSON dump of the seed vectors ------------------------------------------------------------------
namespace {
std::string jstr(const std::string& s) {
    std::string o = "\"";
    for (unsigned char c : s) {
        if (c == '"' || c == '\\') { o += '\\'; o += (char)c; }
        else if (c < 0x20) { char b[8]; snprintf(b, sizeof b, "\\u%04x", c); o += b; }
        else o += (char)c;
    }
    return o + "\"";
}
std::string jnum(double v) {
    if (v != v) return "NaN";
    if (v == INFINITY) return "Infinity";
    if (v == -INFINITY) return "-Infinity";
    char b[40]; snprintf(b, sizeof b, "%.17g", v); return b;
}
std::string jarr(const std::vector<double>& a) { std::string o = "["; for (size_t i = 0; i < a.size(); ++i) { if (i) o += ", "; o += jnum(a[i]); } return o + "]"; }
template <typename T, typename F> std::string jlist(const std::vector<T>& v, F f) {
    std::string o = "[";
    for (size_t i = 0; i < v.size(); ++i) { if (i) o += ", "; o += "[" + f(v[i]) + "]"; }
    return o + "]";
}
}  // namespace

std::string Project::to_json() const {
    const char* T = "true"; const char* F = "false";
    std::string o = "{";
    o += "\"length_seconds\": " + (has_length ? jnum(length_seconds) : std::string("null"));
    o += ", \"render_samplerate\": " + (has_render_sr ? std::to_string(render_samplerate) : std::string("null"));
    o += ", \"render_bitdepth\": " + (has_render_bd ? std::to_string(render_bitdepth) : std::string("null"));
    o += ", \"output_file\": " + (has_output_file ? jstr(output_file) : std::string("null"));
    o += ", \"output_vertex\": " + jstr(output_vertex);
    o += ", \"samples\": " + jlist(samples, [](const SampleSeed& s) { return jstr(s.name) + ", " + jstr(s.path) + ", " + jstr(s.mode); });
    o += ", \"resources\": " + jlist(resources, [](const NamePath& s) { return jstr(s.name) + ", " + jstr(s.path); });
    o += ", \"midis\": " + jlist(midis, [](const NamePath& s) { return jstr(s.name) + ", " + jstr(s.path); });
    o += ", \"streams\": " + jlist(streams, [](const std::string& s) { return jstr(s); });
    o += ", \"lv2plugins\": " + jlist(lv2plugins, [](const NamePath& s) { return jstr(s.name) + ", " + jstr(s.path); });
    o += ", \"lv2params\": " + jlist(lv2params, [](const Lv2Param& s) { return jstr(s.plugin) + ", " + jstr(s.name) + ", " + jnum(s.value); });
    o += ", \"sums\": " + jlist(sums, [](const BaseSeed& s) { return jstr(s.name) + ", " + jnum(s.gain) + ", " + jnum(s.angle); });
    o += ", \"norms\": " + jlist(norms, [](const BaseSeed& s) { return jstr(s.name) + ", " + jnum(s.gain) + ", " + jnum(s.angle); });
    o += ", \"sampleloops\": " + jlist(sampleloops, [](const LoopSeed& s) { return jstr(s.name) + ", " + jnum(s.gain) + ", " + jnum(s.angle) + ", " + jstr(s.sample); });
    o += ", \"samplemultis\": " + jlist(samplemultis, [](const MultiSeed& s) {
        return jstr(s.name) + ", " + jnum(s.gain) + ", " + jnum(s.angle) + ", " + jstr(s.sample) + ", " + jstr(s.floww) + ", " + std::to_string(s.note); });
    o += ", \"samplelerps\": " + jlist(samplelerps, [](const LerpSeed& s) {
        return jstr(s.name) + ", " + jnum(s.gain) + ", " + jnum(s.angle) + ", " + jstr(s.sample) + ", " + jstr(s.floww) + ", " + std::to_string(s.note) + ", " +
               std::to_string(s.lerp_len); });
    o += ", \"debugsines\": " + jlist(debugsines, [](const SineSeed& s) { return jstr(s.name) + ", " + jnum(s.gain) + ", " + jnum(s.angle) + ", " + jstr(s.floww); });
    o += ", \"synths\": " + jlist(synths, [](const SynthSeed& s) {
        return jstr(s.name) + ", " + jnum(s.gain) + ", " + jnum(s.angle) + ", " + jstr(s.floww) + ", " + jnum(s.sq_vel) + ", " + jnum(s.sq_z) + ", " + jarr(s.sq_arr) +
               ", " + jnum(s.tf_vel) + ", " + jnum(s.tf_z) + ", " + jarr(s.tf_arr) + ", " + jnum(s.tr_vel) + ", " + jarr(s.tr_arr); });
    o += ", \"sampsyns\": " + jlist(sampsyns, [](const SampsynSeed& s) {
        return jstr(s.name) + ", " + jnum(s.gain) + ", " + jnum(s.angle) + ", " + jstr(s.floww) + ", " + jarr(s.adsr) + ", " + jstr(s.resource); });
    o += ", \"lv2fxs\": " + jlist(lv2fxs, [](const Lv2fxSeed& s) { return jstr(s.name) + ", " + jnum(s.gain) + ", " + jnum(s.angle) + ", " + jnum(s.wet) + ", " + jstr(s.plugin); });
    o += ", \"adsrs\": " + jlist(adsrs, [T, F](const AdsrSeed& s) {
        return jstr(s.name) + ", " + jnum(s.gain) + ", " + jnum(s.angle) + ", " + jnum(s.wet) + ", " + jstr(s.floww) + ", " + (s.use_off ? T : F) + ", " +
               (s.use_max ? T : F) + ", " + std::to_string(s.note) + ", " + jarr(s.conf); });
    o += ", \"bandpasses\": " + jlist(bandpasses, [T, F](const BandpassSeed& s) {
        return jstr(s.name) + ", " + jnum(s.gain) + ", " + jnum(s.angle) + ", " + jnum(s.wet) + ", " + jnum(s.lo) + ", " + jnum(s.hi) + ", " + (s.pass ? T : F); });
    o += ", \"edges\": " + jlist(edges, [](const std::pair<std::string, std::string>& e) { return jstr(e.first) + ", " + jstr(e.second); });
    return o + "}";
}

}  // namespace tdh
