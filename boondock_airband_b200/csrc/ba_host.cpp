/*
 * ba_host.cpp — libba_host.so: the host-side pieces either side of the CUDA hot path (include/ba_host.h).
 *
 *   conf::      reader for the libconfig grammar subset the reference's configuration files use
 *   model::     the rules of parse_devices()/parse_channels() (src/config.cpp:298-836) producing ba_engine_desc
 *   file input  reader thread with the arithmetic of file_rx_thread / circbuffer_append
 *               (src/input-file.cpp:82-147, src/input-helpers.cpp:37-63)
 *
 * No CUDA here.  Nothing in this file computes a sample: it only decides WHAT the engine is asked to compute.
 */
#include "../../include/ba_host.h"

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cerrno>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <vector>

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

/* carried through the parser and the translator; caught at the C boundary */
struct Problem {
    int code;
    std::string text;
};

[[noreturn]] void raise(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    throw Problem{code, buf};
}

/* ------------------------------------------------------------------------------------------------ grammar */
namespace conf {

enum Kind { K_INT, K_INT64, K_FLOAT, K_BOOL, K_STRING, K_GROUP, K_ARRAY, K_LIST };

struct Node {
    Kind kind = K_GROUP;
    std::string name; /* empty for list/array elements */
    long long i = 0;
    double f = 0;
    bool b = false;
    std::string s;
    std::vector<std::unique_ptr<Node>> kids;
    Node* parent = nullptr;
    int index = -1; /* position in the parent aggregate */
    int line = 0;

    bool aggregate() const { return kind == K_GROUP || kind == K_ARRAY || kind == K_LIST; }
    int length() const { return aggregate() ? (int)kids.size() : 0; } /* Setting::getLength(): 0 for scalars */

    /* Setting::getPath(): dotted names, ".[i]" for unnamed elements */
    std::string path() const {
        std::string p = parent && parent->parent ? parent->path() : "";
        if (!parent)
            return "";
        if (!p.empty())
            p += ".";
        if (name.empty())
            p += "[" + std::to_string(index) + "]";
        else
            p += name;
        return p;
    }

    const Node* find(const char* key) const {
        if (kind != K_GROUP)
            return nullptr;
        for (auto& k : kids)
            if (k->name == key)
                return k.get();
        return nullptr;
    }
    bool exists(const char* key) const { return find(key) != nullptr; }

    /* Setting::operator[](const char*): SettingNotFoundException → "mandatory parameter missing" (.cpp:957-959) */
    const Node& at(const char* key) const {
        const Node* n = find(key);
        if (!n) {
            std::string p = path();
            raise(BA_HOST_ERR_CONFIG, "Configuration error: mandatory parameter missing: %s%s%s", p.c_str(), p.empty() ? "" : ".", key);
        }
        return *n;
    }
    const Node& at(int idx) const {
        if (!aggregate() || idx < 0 || idx >= (int)kids.size()) {
            std::string p = path();
            raise(BA_HOST_ERR_CONFIG, "Configuration error: mandatory parameter missing: %s.[%d]", p.c_str(), idx);
        }
        return *kids[idx];
    }

    /* The conversion operators of libconfig::Setting without auto-conversion (the reference never enables it):
     * a wrong type is a SettingTypeException → "invalid parameter type" (.cpp:960-962). */
    [[noreturn]] void bad_type() const { raise(BA_HOST_ERR_CONFIG, "Configuration error: invalid parameter type: %s", path().c_str()); }
    int as_int() const {
        if (kind == K_INT)
            return (int)i;
        if (kind == K_INT64 && i >= INT32_MIN && i <= INT32_MAX)
            return (int)i;
        bad_type();
    }
    unsigned as_uint() const {
        if ((kind == K_INT || kind == K_INT64) && i >= 0 && i <= (long long)UINT32_MAX)
            return (unsigned)i;
        bad_type();
    }
    double as_double() const {
        if (kind == K_FLOAT)
            return f;
        bad_type();
    }
    float as_float() const { return (float)as_double(); }
    bool as_bool() const {
        if (kind == K_BOOL)
            return b;
        bad_type();
    }
    const char* as_cstr() const {
        if (kind == K_STRING)
            return s.c_str();
        bad_type();
    }
};

class Parser {
  public:
    Parser(const char* text, const std::string& origin, const std::string& dir, int depth) : p_(text), origin_(origin), dir_(dir), depth_(depth) {}

    void parse_settings(Node& group, bool top) {
        for (;;) {
            skip();
            if (!*p_) {
                if (!top)
                    syntax("unexpected end of input inside a group");
                return;
            }
            if (*p_ == '}') {
                if (top)
                    syntax("unmatched '}'");
                return;
            }
            if (*p_ == '@') {
                include(group);
                continue;
            }
            std::string name = ident();
            if (group.find(name.c_str()))
                syntax("duplicate setting name '%s'", name.c_str());
            skip();
            if (*p_ != '=' && *p_ != ':')
                syntax("expected '=' or ':' after '%s'", name.c_str());
            ++p_;
            std::unique_ptr<Node> v = value();
            v->name = name;
            adopt(group, std::move(v));
            skip();
            if (*p_ == ';' || *p_ == ',')
                ++p_;
        }
    }

  private:
    const char* p_;
    std::string origin_, dir_;
    int depth_;
    int line_ = 1;

    [[noreturn]] void syntax(const char* fmt, ...) {
        char buf[512];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof buf, fmt, ap);
        va_end(ap);
        /* the reference prints "Error while parsing configuration file <f> line <n>: <what>" (.cpp:954-956) */
        raise(BA_HOST_ERR_SYNTAX, "Error while parsing configuration file %s line %d: %s", origin_.c_str(), line_, buf);
    }

    static void adopt(Node& parent, std::unique_ptr<Node> kid) {
        kid->parent = &parent;
        kid->index = (int)parent.kids.size();
        parent.kids.push_back(std::move(kid));
    }

    void skip() {
        for (;;) {
            if (*p_ == '\n') {
                ++line_;
                ++p_;
            } else if (*p_ == ' ' || *p_ == '\t' || *p_ == '\r' || *p_ == '\f' || *p_ == '\v') {
                ++p_;
            } else if (*p_ == '#' || (p_[0] == '/' && p_[1] == '/')) {
                while (*p_ && *p_ != '\n')
                    ++p_;
            } else if (p_[0] == '/' && p_[1] == '*') {
                p_ += 2;
                while (*p_ && !(p_[0] == '*' && p_[1] == '/')) {
                    if (*p_ == '\n')
                        ++line_;
                    ++p_;
                }
                if (!*p_)
                    syntax("unterminated comment");
                p_ += 2;
            } else {
                return;
            }
        }
    }

    static bool ident_start(char c) { return (c >= 'A' && c <= 'Z') || (c >= 'a' && c <= 'z') || c == '*' || c == '_'; }
    static bool ident_char(char c) { return ident_start(c) || (c >= '0' && c <= '9') || c == '-'; }

    std::string ident() {
        if (!ident_start(*p_))
            syntax("syntax error near '%.12s'", p_);
        const char* b = p_;
        while (ident_char(*p_))
            ++p_;
        return std::string(b, p_);
    }

    std::string quoted() {
        std::string out;
        for (;;) { /* adjacent string literals concatenate */
            ++p_;  /* opening quote */
            while (*p_ && *p_ != '"') {
                if (*p_ == '\n')
                    ++line_;
                if (*p_ == '\\') {
                    ++p_;
                    switch (*p_) {
                        case 'n': out += '\n'; break;
                        case 'r': out += '\r'; break;
                        case 't': out += '\t'; break;
                        case 'f': out += '\f'; break;
                        case '\\': out += '\\'; break;
                        case '"': out += '"'; break;
                        case 'x': {
                            int v = 0, n = 0;
                            while (n < 2 && isxdigit((unsigned char)p_[1])) {
                                ++p_;
                                v = v * 16 + (isdigit((unsigned char)*p_) ? *p_ - '0' : (tolower(*p_) - 'a' + 10));
                                ++n;
                            }
                            if (!n)
                                syntax("bad \\x escape");
                            out += (char)v;
                            break;
                        }
                        default: syntax("unknown escape in string");
                    }
                    ++p_;
                } else {
                    out += *p_++;
                }
            }
            if (*p_ != '"')
                syntax("unterminated string");
            ++p_;
            skip();
            if (*p_ != '"')
                return out;
        }
    }

    std::unique_ptr<Node> number() {
        std::unique_ptr<Node> n(new Node);
        n->line = line_;
        const char* b = p_;
        if (*p_ == '+' || *p_ == '-')
            ++p_;
        if (p_[0] == '0' && (p_[1] == 'x' || p_[1] == 'X') && isxdigit((unsigned char)p_[2])) {
            char* end;
            unsigned long long v = strtoull(b, &end, 16);
            p_ = end;
            bool wide = false;
            while (*p_ == 'L') {
                wide = true;
                ++p_;
            }
            n->kind = (wide || v > 0xffffffffull) ? K_INT64 : K_INT;
            n->i = n->kind == K_INT ? (long long)(int32_t)(uint32_t)v : (long long)v;
            return n;
        }
        bool digits = false, isfloat = false;
        while (isdigit((unsigned char)*p_)) {
            ++p_;
            digits = true;
        }
        if (*p_ == '.') {
            isfloat = true;
            ++p_;
            while (isdigit((unsigned char)*p_)) {
                ++p_;
                digits = true;
            }
        }
        if (digits && (*p_ == 'e' || *p_ == 'E')) {
            const char* q = p_ + 1;
            if (*q == '+' || *q == '-')
                ++q;
            if (isdigit((unsigned char)*q)) {
                isfloat = true;
                while (isdigit((unsigned char)*q))
                    ++q;
                p_ = q;
            }
        }
        if (!digits)
            syntax("syntax error near '%.12s'", b);
        std::string tok(b, p_);
        if (isfloat) {
            n->kind = K_FLOAT;
            n->f = strtod(tok.c_str(), nullptr);
        } else {
            bool wide = false;
            while (*p_ == 'L') {
                wide = true;
                ++p_;
            }
            n->i = strtoll(tok.c_str(), nullptr, 10);
            n->kind = (wide || n->i > INT32_MAX || n->i < INT32_MIN) ? K_INT64 : K_INT;
        }
        if (ident_char(*p_) && *p_ != '-')
            syntax("syntax error near '%.12s'", b);
        return n;
    }

    std::unique_ptr<Node> value() {
        skip();
        std::unique_ptr<Node> n;
        if (*p_ == '{') {
            n.reset(new Node);
            n->kind = K_GROUP;
            n->line = line_;
            ++p_;
            parse_settings(*n, false);
            ++p_; /* '}' */
        } else if (*p_ == '(' || *p_ == '[') {
            const bool list = *p_ == '(';
            const char close = list ? ')' : ']';
            n.reset(new Node);
            n->kind = list ? K_LIST : K_ARRAY;
            n->line = line_;
            ++p_;
            for (;;) {
                skip();
                if (*p_ == close) {
                    ++p_;
                    break;
                }
                if (!*p_)
                    syntax("unterminated %s", list ? "list" : "array");
                std::unique_ptr<Node> v = value();
                if (!list) {
                    if (v->aggregate())
                        syntax("arrays hold scalar values only");
                    if (!n->kids.empty()) { /* libconfig: all array elements share one type (ints widen to int64) */
                        Kind a = n->kids[0]->kind, bk = v->kind;
                        bool both_int = (a == K_INT || a == K_INT64) && (bk == K_INT || bk == K_INT64);
                        if (a != bk && !both_int)
                            syntax("mismatched element type in array");
                    }
                }
                adopt(*n, std::move(v));
                skip();
                if (*p_ == ',') {
                    ++p_;
                } else if (*p_ != close) {
                    syntax("expected ',' or '%c'", close);
                }
            }
        } else if (*p_ == '"') {
            n.reset(new Node);
            n->kind = K_STRING;
            n->line = line_;
            n->s = quoted();
        } else if (isdigit((unsigned char)*p_) || *p_ == '+' || *p_ == '-' || *p_ == '.') {
            n = number();
        } else if (ident_start(*p_)) {
            const char* b = p_;
            std::string w = ident();
            std::string lw;
            for (char c : w)
                lw += (char)tolower(c);
            if (lw != "true" && lw != "false") {
                p_ = b;
                syntax("syntax error near '%.12s'", b);
            }
            n.reset(new Node);
            n->kind = K_BOOL;
            n->line = line_;
            n->b = lw == "true";
        } else {
            syntax("syntax error near '%.12s'", p_);
        }
        return n;
    }

    void include(Node& group);
};

std::string slurp(const std::string& path) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f)
        raise(BA_HOST_ERR_IO, "Cannot read configuration file %s", path.c_str()); /* .cpp:951-953 */
    std::string s;
    char buf[65536];
    size_t n;
    while ((n = fread(buf, 1, sizeof buf, f)) > 0)
        s.append(buf, n);
    fclose(f);
    return s;
}

void Parser::include(Node& group) {
    const char* b = p_;
    ++p_;
    std::string w = ident();
    if (w != "include") {
        p_ = b;
        syntax("syntax error near '%.12s'", b);
    }
    skip();
    if (*p_ != '"')
        syntax("@include needs a quoted path");
    std::string rel = quoted();
    if (depth_ >= 10)
        syntax("@include nested too deeply");
    std::string path = (rel.empty() || rel[0] == '/' || dir_.empty()) ? rel : dir_ + "/" + rel;
    std::string text = slurp(path);
    size_t slash = path.rfind('/');
    Parser sub(text.c_str(), path, slash == std::string::npos ? "" : path.substr(0, slash), depth_ + 1);
    sub.parse_settings(group, true);
}

}  // namespace conf

/* ------------------------------------------------------------------------------------------------ translation */

using conf::Node;

/* atofs(), util.cpp:130-155: a trailing k/M/G multiplies */
double atofs(const std::string& str) {
    if (str.empty())
        return 0.0;
    double suff = 1.0;
    switch (str.back()) {
        case 'g':
        case 'G': suff *= 1e3; /* fall through */
        case 'm':
        case 'M': suff *= 1e3; /* fall through */
        case 'k':
        case 'K': suff *= 1e3; return suff * atof(str.substr(0, str.size() - 1).c_str());
    }
    return atof(str.c_str());
}

/* parse_anynum2int(), config.cpp:298-310.  Out-of-range doubles would be undefined in the reference's cast; they are
 * saturated here. */
int anynum2int(const Node& n) {
    double v;
    if (n.kind == conf::K_INT)
        return (int)n.i;
    if (n.kind == conf::K_FLOAT)
        v = n.f * 1e6;
    else if (n.kind == conf::K_STRING)
        v = atofs(n.s);
    else
        return 0;
    if (!(v > -2147483649.0))
        return INT32_MIN;
    if (!(v < 2147483648.0))
        return INT32_MAX;
    return (int)v;
}

struct DeviceModel {
    ba_device_desc desc{};
    std::vector<ba_channel_desc> channels;
    std::vector<std::vector<ba_freq_desc>> freq_lists; /* per channel; empty in multichannel mode */
    std::vector<int> source_index;
    std::vector<std::pair<std::string, std::string>> settings;
    bool scan = false;
};

}  // namespace

struct MixerModel {
    std::string name;
    std::vector<ba_mixer_input_desc> inputs;
};

struct ba_conf {
    std::vector<MixerModel> mixers;
    std::vector<ba_mixer_desc> mixer_descs;
    ba_engine_desc desc{};
    std::vector<std::unique_ptr<DeviceModel>> devs;
    std::vector<ba_device_desc> dev_descs;
    std::string warnings;
    int multiple_demod_threads = 0;
};

namespace {

void warn(ba_conf& c, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (!c.warnings.empty())
        c.warnings += "\n";
    c.warnings += buf;
}

bool disabled(const Node& n) { return n.exists("disable") && n.at("disable").as_bool(); }

std::string scalar_text(const Node& n) {
    char buf[64];
    switch (n.kind) {
        case conf::K_INT:
        case conf::K_INT64: snprintf(buf, sizeof buf, "%lld", n.i); return buf;
        case conf::K_FLOAT: snprintf(buf, sizeof buf, "%.17g", n.f); return buf;
        case conf::K_BOOL: return n.b ? "true" : "false";
        case conf::K_STRING: return n.s;
        default: return "";
    }
}

/* the output types parse_outputs() knows (config.cpp:36-265); only what reaches the hot path is kept:
 * a rawfile output sets needs_raw_iq and has_iq_outputs (config.cpp:162) */
int scan_outputs(ba_conf& c, const Node& outs, int i, int j, int dev_index, int chan_index, bool* has_iq) {
    int enabled = 0;
    for (int o = 0; o < outs.length(); o++) {
        const Node& out = outs.at(o);
        if (disabled(out))
            continue;
        const char* type = out.at("type").as_cstr();
        if (!strncmp(type, "icecast", 7) || !strncmp(type, "file", 4) || !strncmp(type, "udp_stream", 6) || !strncmp(type, "pulse", 5)) {
            /* their own mandatory keys belong to the output threads, not to this path */
        } else if (!strncmp(type, "rawfile", 7)) {
            *has_iq = true;
        } else if (!strncmp(type, "mixer", 5)) {
            const char* name = out.at("name").as_cstr();
            MixerModel* mx = nullptr; /* getmixerbyname(): enabled mixers only */
            for (MixerModel& cand : c.mixers)
                if (cand.name == name)
                    mx = &cand;
            if (!mx)
                raise(BA_HOST_ERR_CONFIG, "Configuration error: devices.[%d] channels.[%d] outputs.[%d]: unknown mixer \"%s\"", i, j, o, name);
            const float ampfactor = out.exists("ampfactor") ? out.at("ampfactor").as_float() : 1.0f;
            const float balance = out.exists("balance") ? out.at("balance").as_float() : 0.0f;
            if (balance < -1.0f || balance > 1.0f)
                raise(BA_HOST_ERR_CONFIG, "Configuration error: devices.[%d] channels.[%d] outputs.[%d]: balance out of allowed range <-1.0;1.0>", i, j, o);
            mx->inputs.push_back(ba_mixer_input_desc{dev_index, chan_index, ampfactor, balance}); /* mixer_connect_input, mixer.cpp:55-93 */
        } else {
            raise(BA_HOST_ERR_CONFIG, "Configuration error: devices.[%d] channels.[%d] outputs.[%d]: unknown output type", i, j, o);
        }
        enabled++;
    }
    return enabled;
}

/* parse_channels(), config.cpp:312-729, both modes; R = WAVE_RATE of the build being modelled.  A channel is translated
 * into its frequency list fl[0..freq_count) (one entry in multichannel mode, "freqs" in scan mode) exactly as the
 * reference fills channel->freqlist[f]; the ba_channel_desc carries fl[0] in its own fields and, in scan mode, the list. */
void translate_channels(ba_conf& c, const Node& chans, DeviceModel& dev, int i, int R, bool scan, int fft_size) {
    const bool nfm_build = R == 16000;
    bool slot_needs_raw_iq = false; /* channel_t.needs_raw_iq of slot jj survives a dropped channel (the slot is calloc'ed once) */
    for (int j = 0; j < chans.length(); j++) {
        const Node& ch = chans.at(j);
        if (disabled(ch))
            continue;
        ba_channel_desc d{};
        d.tau_us = -1;
        const int highpass = ch.exists("highpass") ? ch.at("highpass").as_int() : 100;
        const int lowpass = ch.exists("lowpass") ? ch.at("lowpass").as_int() : 2500;
        if (lowpass > 0 && lowpass < highpass)
            raise(BA_HOST_ERR_CONFIG, "Configuration error: devices.[%d] channels.[%d]: lowpass (%d) must be greater than or equal to highpass (%d)", i, j, lowpass, highpass);
        int channel_modulation = BA_MOD_AM;
        if (ch.exists("modulation")) {
            const char* m = ch.at("modulation").as_cstr();
            if (nfm_build && !strncmp(m, "nfm", 3))
                channel_modulation = BA_MOD_NFM;
            else if (strncmp(m, "am", 2) != 0)
                raise(BA_HOST_ERR_CONFIG, "Configuration error: devices.[%d] channels.[%d]: unknown modulation", i, j);
        }
        d.afc = ch.exists("afc") ? (int)(unsigned char)ch.at("afc").as_uint() : 0;
        std::vector<ba_freq_desc> fl;
        auto fresh = [&](int frequency, int modulation) { /* mk_freqlist, config.cpp:268-287 */
            ba_freq_desc f{};
            f.frequency = frequency;
            f.modulation = modulation;
            f.ampfactor = 1.0f;
            f.squelch_snr_threshold = -1.f; /* keep Squelch's default */
            return f;
        };
        if (!scan) {
            fl.push_back(fresh(anynum2int(ch.at("freq")), channel_modulation));
            /* warn_if_freq_not_in_range, config.cpp:289-296 */
            const float bw_limit = (float)dev.desc.sample_rate / 2.f * 0.9f;
            if ((float)abs(fl[0].frequency - dev.desc.centerfreq) >= bw_limit)
                warn(c, "Warning: dev[%d].channel[%d]: frequency %.3f MHz is outside of SDR operating bandwidth (%.3f-%.3f MHz)", i, j, (double)fl[0].frequency / 1e6,
                     (double)(dev.desc.centerfreq - bw_limit) / 1e6, (double)(dev.desc.centerfreq + bw_limit) / 1e6);
            if (ch.exists("label"))
                (void)ch.at("label").as_cstr();
        } else { /* R_SCAN, config.cpp:364-433 */
            const Node& freqs = ch.at("freqs");
            const int n = freqs.length();
            if (n < 1)
                raise(BA_HOST_ERR_CONFIG, "Configuration error: devices.[%d] channels.[%d]: freqs should be a list with at least one element", i, j);
            if (ch.exists("labels") && ch.at("labels").length() < n)
                raise(BA_HOST_ERR_CONFIG, "Configuration error: devices.[%d] channels.[%d]: labels should be a list with at least %d elements", i, j, n);
            struct {
                const char* key;
                const char* what;
            } lists[] = {{"squelch_threshold", "squelch_threshold should be an int or a list of ints"},
                         {"squelch_snr_threshold", "squelch_snr_threshold should be an int, a float or a list of ints or floats"},
                         {"notch", "notch should be an float or a list of floats"},
                         {"notch_q", "notch_q should be a float or a list of floats"},
                         {"ctcss", "ctcss should be an float or a list of floats"}};
            for (auto& l : lists)
                if (ch.exists(l.key) && ch.at(l.key).kind == conf::K_LIST && ch.at(l.key).length() < n)
                    raise(BA_HOST_ERR_CONFIG, "Configuration error: devices.[%d] channels.[%d]: %s with at least %d elements", i, j, l.what, n);
            if (ch.exists("modulation") && ch.exists("modulations"))
                raise(BA_HOST_ERR_CONFIG, "Configuration error: devices.[%d] channels.[%d]: can't set both modulation and modulations", i, j);
            if (ch.exists("modulations") && ch.at("modulations").length() < n)
                raise(BA_HOST_ERR_CONFIG, "Configuration error: devices.[%d] channels.[%d]: modulations should be a list with at least %d elements", i, j, n);
            for (int f = 0; f < n; f++) {
                int mod = channel_modulation;
                if (ch.exists("labels"))
                    (void)ch.at("labels").at(f).as_cstr();
                if (ch.exists("modulations")) {
                    const char* m = ch.at("modulations").at(f).as_cstr();
                    if (nfm_build && !strncmp(m, "nfm", 3))
                        mod = BA_MOD_NFM;
                    else if (!strncmp(m, "am", 2))
                        mod = BA_MOD_AM;
                    else
                        raise(BA_HOST_ERR_CONFIG, "Configuration error: devices.[%d] channels.[%d] modulations.[%d]: unknown modulation", i, j, f);
                }
                fl.push_back(fresh(anynum2int(freqs.at(f)), mod));
            }
            /* "We tune 20 FFT bins higher to avoid DC spike" (config.cpp:429-431): int + 20 * (double)(int / size_t), stored in an int */
            dev.desc.centerfreq = (int)(fl[0].frequency + 20 * (double)((size_t)dev.desc.sample_rate / (size_t)fft_size));
        }
        const int n = (int)fl.size();
        if (ch.exists("squelch"))
            warn(c, "Warning: 'squelch' no longer supported and will be ignored, use 'squelch_threshold' or 'squelch_snr_threshold' instead");
        if (ch.exists("squelch_threshold") && ch.exists("squelch_snr_threshold"))
            warn(c, "Warning: Both 'squelch_threshold' and 'squelch_snr_threshold' are set and may conflict");
        if (ch.exists("squelch_threshold")) {
            const Node& t = ch.at("squelch_threshold");
            if (t.kind != conf::K_LIST && t.kind != conf::K_INT)
                raise(BA_HOST_ERR_CONFIG, "Invalid value for squelch_threshold (should be int or list - use parentheses)");
            for (int f = 0; f < n; f++) {
                const int dbfs = t.kind == conf::K_LIST ? t.at(f).as_int() : (int)t.i;
                if (dbfs > 0)
                    raise(BA_HOST_ERR_CONFIG, "Configuration error: devices.[%d] channels.[%d]: squelch_threshold must be less than or equal to 0", i, j);
                fl[f].squelch_threshold_dbfs = dbfs;
            }
        }
        bool dropped = false;
        if (ch.exists("squelch_snr_threshold")) {
            const Node& t = ch.at("squelch_snr_threshold");
            if (t.kind == conf::K_LIST) {
                for (int f = 0; f < n; f++) {
                    const Node& e = t.at(f);
                    float snr = 0.f;
                    if (e.kind == conf::K_FLOAT)
                        snr = (float)e.f;
                    else if (e.kind == conf::K_INT)
                        snr = (float)(int)e.i;
                    else
                        raise(BA_HOST_ERR_CONFIG, "Configuration error: devices.[%d] channels.[%d]: squelch_snr_threshold list must be of int or float", i, j);
                    if (snr == -1.0f)
                        continue; /* "disable" for this frequency: the default stays */
                    if (snr < 0)
                        raise(BA_HOST_ERR_CONFIG, "Configuration error: devices.[%d] channels.[%d]: squelch_snr_threshold must be greater than or equal to 0", i, j);
                    fl[f].squelch_snr_threshold = snr;
                }
            } else if (t.kind == conf::K_FLOAT || t.kind == conf::K_INT) {
                const float snr = t.kind == conf::K_FLOAT ? (float)t.f : (float)(int)t.i;
                if (snr == -1.0f) {
                    dropped = true; /* `continue` of the channel loop, config.cpp:504-506: the whole channel vanishes */
                } else if (snr < 0) {
                    raise(BA_HOST_ERR_CONFIG, "Configuration error: devices.[%d] channels.[%d]: squelch_snr_threshold must be greater than or equal to 0", i, j);
                } else {
                    for (int f = 0; f < n; f++)
                        fl[f].squelch_snr_threshold = snr;
                }
            } else {
                raise(BA_HOST_ERR_CONFIG, "Invalid value for squelch_snr_threshold (should be float, int, or list of int/float - use parentheses)");
            }
        }
        if (dropped) {
            warn(c, "Note: devices.[%d] channels.[%d] is dropped without a message by the reference (squelch_snr_threshold = -1, config.cpp:504-506)", i, j);
            continue;
        }
        if (ch.exists("notch")) {
            const Node& t = ch.at("notch");
            const Node* q = ch.find("notch_q");
            if (q && q->kind != t.kind)
                raise(BA_HOST_ERR_CONFIG, "Configuration error: devices.[%d] channels.[%d]: notch_q (if set) must be the same type as notch - float or a list of floats with at least %d elements",
                      i, j, n);
            if (t.kind == conf::K_LIST) {
                for (int f = 0; f < n; f++) {
                    const float freq = t.at(f).as_float();
                    float qq = q ? q->at(f).as_float() : 10.0f;
                    if (qq == 0.0f)
                        qq = 10.0f;
                    else if (qq <= 0.0f)
                        raise(BA_HOST_ERR_CONFIG, "Configuration error: devices.[%d] channels.[%d] freq.[%d]: invalid value for notch_q: %g (must be greater than 0.0)", i, j, f, qq);
                    if (freq < 0) {
                        warn(c, "devices.[%d] channels.[%d] freq.[%d]: invalid value for notch: %g, ignoring", i, j, f, freq);
                    } else if (freq > 0) {
                        fl[f].notch = freq;
                        fl[f].notch_q = qq;
                    }
                }
            } else if (t.kind == conf::K_FLOAT) {
                const float freq = (float)t.f;
                const float qq = q ? q->as_float() : 10.0f;
                if (qq <= 0.0f)
                    raise(BA_HOST_ERR_CONFIG, "Configuration error: devices.[%d] channels.[%d]: invalid value for notch_q: %g (must be greater than 0.0)", i, j, qq);
                for (int f = 0; f < n; f++) {
                    if (freq < 0) {
                        warn(c, "devices.[%d] channels.[%d]: notch value '%g' invalid, ignoring", i, j, freq);
                    } else if (freq > 0) {
                        fl[f].notch = freq;
                        fl[f].notch_q = qq;
                    }
                }
            } else {
                raise(BA_HOST_ERR_CONFIG, "Configuration error: devices.[%d] channels.[%d]: notch should be an float or a list of floats with at least %d elements", i, j, n);
            }
        }
        if (ch.exists("ctcss")) {
            const Node& t = ch.at("ctcss");
            if (t.kind != conf::K_LIST && t.kind != conf::K_FLOAT)
                raise(BA_HOST_ERR_CONFIG, "Configuration error: devices.[%d] channels.[%d]: ctcss should be an float or a list of floats with at least %d elements", i, j, n);
            for (int f = 0; f < n; f++) {
                const float freq = t.kind == conf::K_LIST ? t.at(f).as_float() : (float)t.f;
                if (freq < 0 || (freq == 0 && t.kind == conf::K_FLOAT)) {
                    if (t.kind == conf::K_LIST)
                        warn(c, "devices.[%d] channels.[%d] freq.[%d]: invalid value for ctcss: %g, ignoring", i, j, f, freq);
                    else
                        warn(c, "devices.[%d] channels.[%d]: ctcss value '%g' invalid, ignoring", i, j, freq);
                } else if (freq > 0) {
                    fl[f].ctcss = freq;
                }
            }
        }
        bool needs_raw_iq = slot_needs_raw_iq;
        if (ch.exists("bandwidth")) {
            needs_raw_iq = slot_needs_raw_iq = true;
            const Node& t = ch.at("bandwidth");
            if (t.kind == conf::K_LIST) {
                for (int f = 0; f < n; f++) {
                    const int bw = anynum2int(t.at(f));
                    if (bw < 0)
                        warn(c, "devices.[%d] channels.[%d] freq.[%d]: bandwidth value '%d' invalid, ignoring", i, j, f, bw);
                    else
                        fl[f].bandwidth = bw; /* 0: "disable" for this frequency */
                }
            } else {
                const int bw = anynum2int(t);
                if (bw == 0) {
                    warn(c, "Note: devices.[%d] channels.[%d] is dropped without a message by the reference (bandwidth = 0, config.cpp:612-614)", i, j);
                    continue;
                }
                if (bw < 0)
                    warn(c, "devices.[%d] channels.[%d]: bandwidth value '%d' invalid, ignoring", i, j, bw);
                else
                    for (int f = 0; f < n; f++)
                        fl[f].bandwidth = bw;
            }
        }
        if (ch.exists("ampfactor")) {
            const Node& t = ch.at("ampfactor");
            for (int f = 0; f < n; f++) {
                const float a = t.kind == conf::K_LIST ? t.at(f).as_float() : t.as_float();
                if (a < 0)
                    raise(BA_HOST_ERR_CONFIG, "devices.[%d] channels.[%d]: ampfactor '%g' must not be negative", i, j, a);
                fl[f].ampfactor = a;
            }
        }
        if (nfm_build && ch.exists("tau"))
            d.tau_us = ch.at("tau").as_int();
        const Node& outs = ch.at("outputs");
        bool has_iq = false;
        if (outs.length() < 1 || scan_outputs(c, outs, i, j, (int)c.devs.size(), (int)dev.channels.size(), &has_iq) < 1)
            raise(BA_HOST_ERR_CONFIG, "Configuration error: devices.[%d] channels.[%d]: no outputs defined", i, j);
        d.has_iq_outputs = has_iq ? 1 : 0;
        /* needs_raw_iq without a low-pass anywhere (bandwidth present but rejected or zero in every list entry, or inherited from a
         * dropped entry's slot): expressed as bandwidth < 0 on the first frequency, which the engine reads as "raw IQ path on, no filter" */
        bool implied = has_iq;
        for (const ba_freq_desc& f : fl)
            implied = implied || f.modulation == BA_MOD_NFM || f.bandwidth != 0;
        if (needs_raw_iq && !implied)
            fl[0].bandwidth = -1;
        slot_needs_raw_iq = false; /* the next slot is fresh */
        d.frequency = fl[0].frequency;
        d.modulation = fl[0].modulation;
        d.ampfactor = fl[0].ampfactor;
        d.squelch_threshold_dbfs = fl[0].squelch_threshold_dbfs;
        d.squelch_snr_threshold = fl[0].squelch_snr_threshold;
        d.notch = fl[0].notch;
        d.notch_q = fl[0].notch_q;
        d.ctcss = fl[0].ctcss;
        d.bandwidth = fl[0].bandwidth;
        if (scan) {
            dev.freq_lists.push_back(fl); /* pointers are set once the device is complete */
            d.freq_count = n;
        } else {
            dev.freq_lists.emplace_back();
        }
        dev.channels.push_back(d);
        dev.source_index.push_back(j);
    }
}

int sample_format_by_name(const char* s, int* bytes, float* fullscale) {
    std::string u;
    for (const char* p = s; *p; ++p)
        u += (char)toupper(*p);
    if (u == "U8" || u == "CU8") {
        *bytes = 1;
        *fullscale = 126.5f; /* (float)SCHAR_MAX - 0.5f */
        return BA_SFMT_U8;
    }
    if (u == "S8" || u == "CS8") {
        *bytes = 1;
        *fullscale = 126.5f; /* (float)SCHAR_MAX - 0.5f */
        return BA_SFMT_S8;
    }
    if (u == "S16" || u == "CS16") {
        *bytes = 2;
        *fullscale = 32766.5f;
        return BA_SFMT_S16;
    }
    if (u == "F32" || u == "CF32") {
        *bytes = 4;
        *fullscale = 1.0f;
        return BA_SFMT_F32;
    }
    return BA_SFMT_UNDEF;
}

/* parse_devices(), config.cpp:731-836, and the root keys main() reads before it (.cpp:846-893) */
void translate(ba_conf& c, const Node& root, int wave_rate) {
    int R = wave_rate;
    if (R == 0) { /* pick the build from the file: any "nfm" modulation needs the NFM build */
        R = 8000;
        if (root.exists("devices")) {
            const Node& devs = root.at("devices");
            for (int i = 0; i < devs.length() && R == 8000; i++) {
                const Node* chans = devs.at(i).find("channels");
                for (int j = 0; chans && j < chans->length(); j++) {
                    const Node* m = chans->at(j).find("modulation");
                    const Node* ms = chans->at(j).find("modulations");
                    if (m && m->kind == conf::K_STRING && !strncmp(m->s.c_str(), "nfm", 3))
                        R = 16000;
                    for (int f = 0; ms && f < ms->length(); f++)
                        if (ms->at(f).kind == conf::K_STRING && !strncmp(ms->at(f).s.c_str(), "nfm", 3))
                            R = 16000;
                }
            }
        }
    }
    if (R != 8000 && R != 16000)
        raise(BA_ERR_BAD_ARG, "wave_rate must be 8000, 16000 or 0");
    const bool nfm_build = R == 16000;
    int fft_size = 512; /* DEFAULT_FFT_SIZE_LOG 9, boondock_airband.h:81 */
    if (root.exists("fft_size")) {
        const int fsize = root.at("fft_size").as_int();
        bool ok = false;
        for (int lg = 8; lg <= 13; lg++)
            ok |= fsize == 1 << lg;
        if (!ok)
            raise(BA_HOST_ERR_CONFIG, "Configuration error: invalid fft_size value (must be a power of two in range 256-8192)");
        fft_size = fsize;
    }
    c.multiple_demod_threads = root.exists("multiple_demod_threads") && root.at("multiple_demod_threads").as_bool();
    int global_tau = -1; /* alpha = exp(-1/(WAVE_RATE*2e-4)) unless root.tau (.cpp:87,891-893) */
    if (nfm_build && root.exists("tau"))
        global_tau = root.at("tau").as_int();
    const Node& devs = root.at("devices");
    if (devs.length() < 1)
        raise(BA_HOST_ERR_CONFIG, "Configuration error: no devices defined");
    /* parse_mixers(), config.cpp:838-889: runs before the devices so that channel outputs can name a mixer */
    if (const Node* mixers = root.find("mixers")) {
        for (int i = 0; i < mixers->length(); i++) {
            const Node& mn = mixers->at(i);
            if (mn.kind != conf::K_GROUP)
                mn.bad_type();
            if (disabled(mn))
                continue;
            if (mn.name.empty())
                raise(BA_HOST_ERR_CONFIG, "Configuration error: mixers.[%d]: undefined mixer name", i);
            const int highpass = mn.exists("highpass") ? mn.at("highpass").as_int() : 100;
            const int lowpass = mn.exists("lowpass") ? mn.at("lowpass").as_int() : 2500;
            if (lowpass > 0 && lowpass < highpass)
                raise(BA_HOST_ERR_CONFIG, "Configuration error: mixers.[%d]: lowpass (%d) must be greater than or equal to highpass (%d)", i, lowpass, highpass);
            const Node& outs = mn.at("outputs");
            int enabled = 0;
            for (int o = 0; o < outs.length(); o++) {
                const Node& out = outs.at(o);
                if (disabled(out))
                    continue;
                const char* type = out.at("type").as_cstr();
                if (!strncmp(type, "rawfile", 7))
                    raise(BA_HOST_ERR_CONFIG, "Configuration error: mixers.[%d] outputs[%d]: rawfile output is not allowed for mixers", i, o);
                if (!strncmp(type, "mixer", 5))
                    raise(BA_HOST_ERR_CONFIG, "Configuration error: mixers.[%d] outputs.[%d]: mixer output is not allowed for mixers", i, o);
                if (strncmp(type, "icecast", 7) && strncmp(type, "file", 4) && strncmp(type, "udp_stream", 6) && strncmp(type, "pulse", 5))
                    raise(BA_HOST_ERR_CONFIG, "Configuration error: mixers.[%d] outputs.[%d]: unknown output type", i, o);
                enabled++;
            }
            if (enabled < 1)
                raise(BA_HOST_ERR_CONFIG, "Configuration error: mixers.[%d]: no outputs defined", i);
            c.mixers.push_back(MixerModel{mn.name, {}});
        }
    }
    for (int i = 0; i < devs.length(); i++) {
        const Node& dn = devs.at(i);
        if (disabled(dn))
            continue;
        std::unique_ptr<DeviceModel> dev(new DeviceModel);
        ba_device_desc& d = dev->desc;
        std::string type;
        if (dn.exists("type")) {
            type = dn.at("type").as_cstr();
        } else {
            warn(c, "Warning: devices.[%d]: assuming device type \"rtlsdr\", please set \"type\" in the device section.", i);
            type = "rtlsdr";
        }
        /* what <type>_input_new() presets (input-rtlsdr.cpp:244-247, input-mirisdr.cpp:229-232, input-file.cpp:168-173,
         * input-soapysdr.cpp:355-358) */
        d.sample_rate = 0;
        if (type == "rtlsdr" || type == "file") {
            d.sample_format = BA_SFMT_U8;
            d.bytes_per_sample = 1;
            d.fullscale = 126.5f; /* (float)SCHAR_MAX - 0.5f */
            if (type == "rtlsdr")
                d.sample_rate = 2560000;
        } else if (type == "mirisdr") {
            d.sample_format = BA_SFMT_S8;
            d.bytes_per_sample = 1;
            d.fullscale = 126.5f; /* (float)SCHAR_MAX - 0.5f */
            d.sample_rate = 2560000;
        } else if (type == "soapysdr") {
            d.sample_format = BA_SFMT_UNDEF; /* known only once the device is opened, unless "sample_format" says */
            d.sample_rate = -1;
        } else {
            raise(BA_HOST_ERR_CONFIG, "Configuration error: devices.[%d]: unsupported device type", i);
        }
        if (dn.exists("sample_rate")) {
            const int sr = anynum2int(dn.at("sample_rate"));
            if (sr < R)
                raise(BA_HOST_ERR_CONFIG, "Configuration error: devices.[%d]: sample_rate must be greater than %d", i, R);
            d.sample_rate = sr;
        }
        bool scan = false; /* R_SCAN / R_MULTICHANNEL, config.cpp:761-772 */
        if (dn.exists("mode")) {
            const char* m = dn.at("mode").as_cstr();
            if (!strncmp(m, "multichannel", 12)) {
            } else if (!strncmp(m, "scan", 4)) {
                scan = true;
            } else {
                raise(BA_HOST_ERR_CONFIG, "Configuration error: devices.[%d]: invalid mode (must be one of: \"scan\", \"multichannel\")", i);
            }
        }
        if (!scan) /* in scan mode parse_channels() derives it from the first frequency (config.cpp:773-775,429-431) */
            d.centerfreq = anynum2int(dn.at("centerfreq"));
        d.tau_us = global_tau;
        if (nfm_build && dn.exists("tau"))
            d.tau_us = dn.at("tau").as_int();
        /* driver settings: kept as text for the driver; the file driver's own checks (input-file.cpp:35-66) run here */
        for (auto& k : dn.kids)
            if (!k->aggregate())
                dev->settings.emplace_back(k->name, scalar_text(*k));
        if (type == "file") {
            if (!dn.exists("filepath"))
                raise(BA_HOST_ERR_CONFIG, "File configuration error: no 'filepath' given");
            (void)dn.at("filepath").as_cstr();
            if (dn.exists("speedup_factor")) {
                const Node& s = dn.at("speedup_factor");
                double v;
                if (s.kind == conf::K_INT)
                    v = (double)s.i;
                else if (s.kind == conf::K_FLOAT)
                    v = (float)s.f;
                else
                    raise(BA_HOST_ERR_CONFIG, "File configuration error: 'speedup_factor' must be a float or int if set");
                if (v <= 0.0)
                    raise(BA_HOST_ERR_CONFIG, "File configuration error: 'speedup_factor' must be >= 0.0");
            }
        }
        /* extension of this framework (row f-1): a file or SoapySDR input may name its sample format */
        if (dn.exists("sample_format")) {
            int bytes = 0;
            float fs = 0;
            const int f = sample_format_by_name(dn.at("sample_format").as_cstr(), &bytes, &fs);
            if (f == BA_SFMT_UNDEF)
                raise(BA_HOST_ERR_CONFIG, "Configuration error: devices.[%d]: sample_format must be one of U8, S8, S16, F32", i);
            if (type != "file" && type != "soapysdr")
                raise(BA_HOST_ERR_CONFIG, "Configuration error: devices.[%d]: sample_format applies to file and soapysdr inputs only", i);
            d.sample_format = f;
            d.bytes_per_sample = bytes;
            d.fullscale = fs;
        }
        if (dn.exists("fullscale")) {
            const Node& s = dn.at("fullscale");
            d.fullscale = s.kind == conf::K_INT ? (float)s.i : s.as_float();
        }
        /* the assertions of config.cpp:790-793 */
        if (d.sample_format == BA_SFMT_UNDEF)
            raise(BA_HOST_ERR_CONFIG, "Configuration error: devices.[%d]: the sample format of a soapysdr input is known only once the device is opened; set \"sample_format\"", i);
        if (!(d.fullscale > 0))
            raise(BA_HOST_ERR_CONFIG, "Configuration error: devices.[%d]: fullscale must be positive", i);
        if (d.sample_rate <= R)
            raise(BA_HOST_ERR_CONFIG, "Configuration error: devices.[%d]: sample_rate must be greater than %d", i, R);
        const Node& chans = dn.at("channels");
        if (chans.length() < 1)
            raise(BA_HOST_ERR_CONFIG, "Configuration error: devices.[%d]: no channels configured", i);
        translate_channels(c, chans, *dev, i, R, scan, fft_size);
        if (dev->channels.empty())
            raise(BA_HOST_ERR_CONFIG, "Configuration error: devices.[%d]: no channels enabled", i);
        if (scan && dev->channels.size() > 1)
            raise(BA_HOST_ERR_CONFIG, "Configuration error: devices.[%d]: only one channel is allowed in scan mode", i);
        dev->scan = scan;
        c.devs.push_back(std::move(dev));
    }
    if (c.devs.empty())
        raise(BA_HOST_ERR_CONFIG, "Configuration error: no devices defined");
    for (auto& dev : c.devs) {
        for (size_t k = 0; k < dev->channels.size(); k++)
            dev->channels[k].freqs = dev->freq_lists[k].empty() ? nullptr : dev->freq_lists[k].data();
        dev->desc.channel_count = (int32_t)dev->channels.size();
        dev->desc.channels = dev->channels.data();
        c.dev_descs.push_back(dev->desc);
    }
    c.desc.abi_version = BA_CUDA_ABI_VERSION;
    c.desc.fft_size = fft_size;
    c.desc.wave_rate = R;
    c.desc.fm_demod = BA_FM_FAST_ATAN2;
    c.desc.device_count = (int32_t)c.dev_descs.size();
    c.desc.devices = c.dev_descs.data();
    for (MixerModel& mx : c.mixers)
        c.mixer_descs.push_back(ba_mixer_desc{(int32_t)mx.inputs.size(), mx.inputs.data()});
    c.desc.mixer_count = (int32_t)c.mixer_descs.size();
    c.desc.mixers = c.mixer_descs.data();
}

int parse_into(const char* text, const std::string& origin, const std::string& dir, int wave_rate, ba_conf** out) {
    if (!text || !out)
        return fail(BA_ERR_BAD_ARG, "bad argument");
    *out = nullptr;
    std::unique_ptr<ba_conf> c(new (std::nothrow) ba_conf);
    if (!c)
        return fail(BA_ERR_NOMEM, "out of memory");
    try {
        Node root;
        conf::Parser p(text, origin, dir, 0);
        p.parse_settings(root, true);
        translate(*c, root, wave_rate);
    } catch (const Problem& pr) {
        g_err = pr.text;
        return pr.code;
    } catch (const std::bad_alloc&) {
        return fail(BA_ERR_NOMEM, "out of memory");
    }
    *out = c.release();
    return BA_OK;
}

}  // namespace

extern "C" {

const char* ba_host_last_error(void) { return g_err.c_str(); }

int ba_conf_parse_text(const char* text, int wave_rate, ba_conf** out) { return parse_into(text, "<text>", "", wave_rate, out); }

int ba_conf_parse_file(const char* path, int wave_rate, ba_conf** out) {
    if (!path || !out)
        return fail(BA_ERR_BAD_ARG, "bad argument");
    std::string text;
    try {
        text = conf::slurp(path);
    } catch (const Problem& pr) {
        g_err = pr.text;
        return pr.code;
    }
    std::string p(path);
    size_t slash = p.rfind('/');
    return parse_into(text.c_str(), p, slash == std::string::npos ? "" : p.substr(0, slash), wave_rate, out);
}

void ba_conf_free(ba_conf* c) { delete c; }

const ba_engine_desc* ba_conf_engine_desc(const ba_conf* c) { return c ? &c->desc : nullptr; }
int ba_conf_device_count(const ba_conf* c) { return c ? (int)c->devs.size() : 0; }

const char* ba_conf_device_setting(const ba_conf* c, int device, const char* key) {
    if (!c || !key || device < 0 || device >= (int)c->devs.size())
        return nullptr;
    for (auto& kv : c->devs[device]->settings)
        if (kv.first == key)
            return kv.second.c_str();
    return nullptr;
}

const char* ba_conf_mixer_name(const ba_conf* c, int mixer) {
    return c && mixer >= 0 && mixer < (int)c->mixers.size() ? c->mixers[mixer].name.c_str() : nullptr;
}

int ba_conf_device_is_scan(const ba_conf* c, int device) {
    return c && device >= 0 && device < (int)c->devs.size() && c->devs[device]->scan ? 1 : 0;
}

/* controller_thread(), boondock_airband.cpp:101-139, as a step function: one call per 200 ms poll.  state = {i,
 * consecutive_squelch_off, last_frequency, freq_count}.  Returns the freq_idx the channel is on after the poll (hand it to
 * ba_cuda_set_freq_idx and retune the input when it changed); *tag_freq (optional) is set to i when the reference would
 * queue a tag for the outputs' metadata (squelch just opened on a new frequency), else -1. */
int ba_scan_controller_poll(int32_t state[4], int has_signal, int* tag_freq) {
    if (tag_freq)
        *tag_freq = -1;
    if (!state || state[3] < 2)
        return state ? state[0] : 0;
    if (!has_signal) { /* axcindicate == NO_SIGNAL */
        if (state[1] < 10) {
            state[1]++;
        } else {
            state[0] = (state[0] + 1) % state[3];
        }
    } else {
        if (state[1] == 10 && state[0] != state[2]) {
            if (tag_freq)
                *tag_freq = state[0];
            state[2] = state[0];
        }
        state[1] = 0;
    }
    return state[0];
}

int ba_conf_multiple_demod_threads(const ba_conf* c) { return c ? c->multiple_demod_threads : 0; }
const char* ba_conf_warnings(const ba_conf* c) { return c ? c->warnings.c_str() : ""; }

int ba_conf_channel_source_index(const ba_conf* c, int device, int ch) {
    if (!c || device < 0 || device >= (int)c->devs.size())
        return -1;
    auto& v = c->devs[device]->source_index;
    return ch >= 0 && ch < (int)v.size() ? v[ch] : -1;
}

}  // extern "C"

/* ------------------------------------------------------------------------------------------------ file input */

struct ba_file_input {
    ba_file_input_desc desc{};
    std::string path;
    ba_ring_sink sink{};
    FILE* fp = nullptr;
    std::thread th;
    std::atomic<int> state{BA_INPUT_UNKNOWN};
    std::atomic<bool> quit{false};
    std::atomic<uint64_t> bytes{0};
    bool started = false;
};

namespace {

struct EngineSink {
    ba_engine* e;
    int dev;
    ba_submit_fn submit;
    ba_space_fn space;
};

size_t engine_space(void* ctx) {
    EngineSink* s = (EngineSink*)ctx;
    size_t n = 0;
    return s->space(s->e, s->dev, &n) == BA_OK ? n : 0;
}
int engine_append(void* ctx, const void* data, size_t n) {
    EngineSink* s = (EngineSink*)ctx;
    return s->submit(s->e, s->dev, data, n);
}

/* file_rx_thread, input-file.cpp:82-147.  Differences, both asked for by row f-1: any sample format (whole complex
 * samples are kept together), and a full ring is waited for in short naps instead of a 10 ms sleep per poll. */
void reader(ba_file_input* f) {
    const size_t unit = 2 * (size_t)(f->desc.sample_format == BA_SFMT_S16 ? 2 : f->desc.sample_format == BA_SFMT_F32 ? 4 : 1);
    size_t buf_len = f->desc.chunk_bytes ? f->desc.chunk_bytes : (f->desc.ring_bytes / 2) - 1; /* input-file.cpp:96 */
    buf_len -= buf_len % unit;
    if (buf_len == 0)
        buf_len = unit;
    std::vector<unsigned char> buf(buf_len);
    /* 1000 / (sample_rate * bytes_per_sample * 2 * speedup_factor), input-file.cpp:99 */
    const double ms_per_byte = f->desc.speedup_factor > 0 ? 1000.0 / ((double)f->desc.sample_rate * (unit / 2) * 2 * f->desc.speedup_factor) : 0.0;
    f->state = BA_INPUT_RUNNING;
    int nap_us = 20;
    while (!f->quit) {
        if (feof(f->fp)) {
            if (f->desc.loop) {
                clearerr(f->fp);
                if (fseek(f->fp, 0, SEEK_SET) != 0) {
                    f->state = BA_INPUT_FAILED;
                    break;
                }
            } else {
                f->state = BA_INPUT_FAILED; /* "hit end of file, disabling", input-file.cpp:107-111 */
                break;
            }
        }
        if (ferror(f->fp)) {
            f->state = BA_INPUT_FAILED;
            break;
        }
        const auto t0 = std::chrono::steady_clock::now();
        if (f->sink.space(f->sink.ctx) >= buf_len) {
            nap_us = 20;
            const size_t len = fread(buf.data(), 1, buf_len, f->fp);
            const size_t whole = len - len % unit; /* a trailing partial sample at end of file is dropped */
            if (whole) {
                if (f->sink.append(f->sink.ctx, buf.data(), whole) != 0) {
                    f->state = BA_INPUT_FAILED;
                    break;
                }
                f->bytes += whole;
            }
            if (ms_per_byte > 0) {
                const double took = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
                const int sleep_ms = (int)(len * ms_per_byte - (int)took);
                if (sleep_ms > 0)
                    std::this_thread::sleep_for(std::chrono::milliseconds(sleep_ms));
            }
        } else {
            std::this_thread::sleep_for(std::chrono::microseconds(nap_us));
            if (nap_us < 1000)
                nap_us *= 2;
        }
    }
    if (f->quit && f->state == BA_INPUT_RUNNING)
        f->state = BA_INPUT_STOPPED;
}

}  // namespace

extern "C" {

int ba_file_input_sink_for_engine(ba_engine* e, int dev, ba_submit_fn submit, ba_space_fn space, ba_ring_sink* out) {
    if (!e || !submit || !space || !out)
        return fail(BA_ERR_BAD_ARG, "bad argument");
    EngineSink* s = new (std::nothrow) EngineSink{e, dev, submit, space};
    if (!s)
        return fail(BA_ERR_NOMEM, "out of memory");
    out->ctx = s; /* a few bytes, owned by the sink for the life of the process side of the input */
    out->space = engine_space;
    out->append = engine_append;
    return BA_OK;
}

void ba_file_input_sink_release(ba_ring_sink* sink) {
    if (sink && sink->space == engine_space) {
        delete (EngineSink*)sink->ctx;
        sink->ctx = nullptr;
        sink->space = nullptr;
        sink->append = nullptr;
    }
}

int ba_file_input_open(const ba_file_input_desc* desc, const ba_ring_sink* sink, ba_file_input** out) {
    if (!desc || !sink || !out || !desc->filepath || !sink->space || !sink->append)
        return fail(BA_ERR_BAD_ARG, "bad argument");
    *out = nullptr;
    if (desc->sample_format < BA_SFMT_U8 || desc->sample_format > BA_SFMT_F32)
        return fail(BA_ERR_BAD_ARG, "unknown sample format %d", desc->sample_format);
    if (desc->speedup_factor < 0)
        return fail(BA_HOST_ERR_CONFIG, "File configuration error: 'speedup_factor' must be >= 0.0");
    if (desc->speedup_factor > 0 && desc->sample_rate <= 0)
        return fail(BA_ERR_BAD_ARG, "a paced replay needs the sample rate");
    if (!desc->chunk_bytes && desc->ring_bytes < 16)
        return fail(BA_ERR_BAD_ARG, "neither chunk_bytes nor ring_bytes given");
    std::unique_ptr<ba_file_input> f(new (std::nothrow) ba_file_input);
    if (!f)
        return fail(BA_ERR_NOMEM, "out of memory");
    f->desc = *desc;
    f->path = desc->filepath;
    f->desc.filepath = f->path.c_str();
    f->sink = *sink;
    f->fp = fopen(f->path.c_str(), "rb");
    if (!f->fp) /* file_init: the reference exits here (input-file.cpp:73-77) */
        return fail(BA_HOST_ERR_IO, "File input failed to open '%s' - %s", f->path.c_str(), strerror(errno));
    f->state = BA_INPUT_INITIALIZED;
    *out = f.release();
    return BA_OK;
}

int ba_file_input_start(ba_file_input* f) {
    if (!f || f->started)
        return fail(BA_ERR_STATE, "file input not open or already started");
    f->started = true;
    try {
        f->th = std::thread(reader, f);
    } catch (...) {
        f->started = false;
        f->state = BA_INPUT_FAILED;
        return fail(BA_ERR_NOMEM, "cannot start the reader thread");
    }
    return BA_OK;
}

int ba_file_input_state(const ba_file_input* f) { return f ? f->state.load() : BA_INPUT_UNKNOWN; }
uint64_t ba_file_input_bytes(const ba_file_input* f) { return f ? f->bytes.load() : 0; }

int ba_file_input_stop(ba_file_input* f) {
    if (!f)
        return BA_OK;
    f->quit = true;
    if (f->th.joinable())
        f->th.join();
    if (f->fp)
        fclose(f->fp);
    delete f;
    return BA_OK;
}

}  // extern "C"

/* ------------------------------------------------------------------------------------------------ output hand-off */

struct ba_handoff {
    std::mutex lock;
    std::condition_variable can_fill, can_take;
    unsigned char* arena = nullptr;
    size_t slot_bytes = 0, pitch = 0;
    int slots = 0;
    std::vector<int> free_list;                   /* slot indices nobody holds */
    std::deque<std::pair<int, uint64_t>> ready;   /* published, oldest first: (slot, tag) */
    std::vector<uint8_t> state;                   /* 0 free, 1 being filled, 2 published, 3 being read */
    bool closed = false;
    std::atomic<uint64_t> overruns{0};
};

namespace {

int slot_index(const ba_handoff* h, const void* slot) {
    const unsigned char* p = (const unsigned char*)slot;
    if (!h || !slot || p < h->arena || p >= h->arena + h->pitch * h->slots || (size_t)(p - h->arena) % h->pitch)
        return -1;
    return (int)((size_t)(p - h->arena) / h->pitch);
}

template <class Pred>
bool wait_on(std::condition_variable& cv, std::unique_lock<std::mutex>& g, int timeout_ms, Pred pred) {
    if (timeout_ms < 0) {
        cv.wait(g, pred);
        return true;
    }
    return cv.wait_for(g, std::chrono::milliseconds(timeout_ms), pred);
}

}  // namespace

extern "C" {

int ba_handoff_create(int slots, size_t slot_bytes, ba_handoff** out) {
    if (slots < 1 || slot_bytes == 0 || !out)
        return fail(BA_ERR_BAD_ARG, "bad argument");
    *out = nullptr;
    std::unique_ptr<ba_handoff> h(new (std::nothrow) ba_handoff);
    if (!h)
        return fail(BA_ERR_NOMEM, "out of memory");
    h->pitch = (slot_bytes + 63) & ~(size_t)63;
    h->slot_bytes = slot_bytes;
    h->slots = slots;
    h->arena = (unsigned char*)aligned_alloc(64, h->pitch * (size_t)slots);
    if (!h->arena)
        return fail(BA_ERR_NOMEM, "%d hand-off slots of %zu bytes", slots, slot_bytes);
    h->state.assign(slots, 0);
    for (int i = slots - 1; i >= 0; i--)
        h->free_list.push_back(i);
    *out = h.release();
    return BA_OK;
}

int ba_handoff_acquire(ba_handoff* h, int timeout_ms, void** slot) {
    if (!h || !slot)
        return fail(BA_ERR_BAD_ARG, "bad argument");
    std::unique_lock<std::mutex> g(h->lock);
    const bool ok = wait_on(h->can_fill, g, timeout_ms, [&] { return h->closed || !h->free_list.empty(); });
    if (h->closed)
        return BA_HANDOFF_CLOSED;
    if (!ok) {
        h->overruns++; /* the consumer has not kept up: output_overrun_count++, boondock_airband.cpp:673-676 */
        return BA_HANDOFF_TIMEOUT;
    }
    const int i = h->free_list.back();
    h->free_list.pop_back();
    h->state[i] = 1;
    *slot = h->arena + h->pitch * (size_t)i;
    return BA_OK;
}

int ba_handoff_publish(ba_handoff* h, void* slot, uint64_t tag) {
    const int i = slot_index(h, slot);
    if (i < 0)
        return fail(BA_ERR_BAD_ARG, "not a slot of this hand-off");
    {
        std::lock_guard<std::mutex> g(h->lock);
        if (h->state[i] != 1)
            return fail(BA_ERR_STATE, "slot %d was not acquired", i);
        h->state[i] = 2;
        h->ready.emplace_back(i, tag);
    }
    h->can_take.notify_one();
    return BA_OK;
}

int ba_handoff_take(ba_handoff* h, int timeout_ms, void** slot, uint64_t* tag) {
    if (!h || !slot)
        return fail(BA_ERR_BAD_ARG, "bad argument");
    std::unique_lock<std::mutex> g(h->lock);
    const bool ok = wait_on(h->can_take, g, timeout_ms, [&] { return h->closed || !h->ready.empty(); });
    if (h->ready.empty())
        return h->closed ? BA_HANDOFF_CLOSED : (ok ? BA_HANDOFF_CLOSED : BA_HANDOFF_TIMEOUT);
    const std::pair<int, uint64_t> r = h->ready.front();
    h->ready.pop_front();
    h->state[r.first] = 3;
    *slot = h->arena + h->pitch * (size_t)r.first;
    if (tag)
        *tag = r.second;
    return BA_OK;
}

int ba_handoff_release(ba_handoff* h, void* slot) {
    const int i = slot_index(h, slot);
    if (i < 0)
        return fail(BA_ERR_BAD_ARG, "not a slot of this hand-off");
    {
        std::lock_guard<std::mutex> g(h->lock);
        if (h->state[i] != 3)
            return fail(BA_ERR_STATE, "slot %d was not taken", i);
        h->state[i] = 0;
        h->free_list.push_back(i);
    }
    h->can_fill.notify_one();
    return BA_OK;
}

void ba_handoff_close(ba_handoff* h) {
    if (!h)
        return;
    {
        std::lock_guard<std::mutex> g(h->lock);
        h->closed = true;
    }
    h->can_fill.notify_all();
    h->can_take.notify_all();
}

uint64_t ba_handoff_overruns(const ba_handoff* h) { return h ? h->overruns.load() : 0; }

void ba_handoff_destroy(ba_handoff* h) {
    if (!h)
        return;
    free(h->arena);
    delete h;
}

}  // extern "C"
