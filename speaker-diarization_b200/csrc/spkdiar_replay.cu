// spkdiar_replay.cu - host-only translation unit: the text side of the two stages spk-diarization2.py:122-128
// runs per recording (`spk-change-detection.py -m gw ...` then `spk-clustering.py -m hi ...`), as native code.
//
// The device returns window records (gw.cuh) and merge sequences (cluster*.cuh); turning them into recipe text is
// what the reference scripts do around their numeric kernels:
//   recipe parsing                     spk-change-detection.py:11-28 (four independent regular-expression searches)
//   one chain per run of equal lna     spk-change-detection.py:370-374
//   a recipe line per detected change  spk-change-detection.py:252-256, 286-288 -> write_recipe_line, 46-69
//   the clustering stage reads that recipe back (text!), builds one cluster per line   spk-clustering.py:263-283
//   merges applied to the speaker lists, turns written smallest first                  spk-clustering.py:216-220, 243-260
// with Python-2 number text (str(float) = '%.12g', '.0' appended) and the LNA renaming state of write_recipe_line.
// The Python classes Detector / Clusterer do the same (and everything else: logs, other modes, other flags); this
// unit covers exactly the corpus configuration and answers SPKDIAR_E_UNSUPPORTED for anything it does not
// reproduce to the byte (non-ASCII text, exponents or signs in time fields, several wavs in one recipe), so that
// the caller can hand that recording to the Python replay.  No device, no context: plain C++.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include <algorithm>
#include <new>
#include <string>
#include <vector>

#include "../../include/spkdiar.h"

namespace {

struct Line { std::string audio, lna; double start, end; };

struct LnaState {                                  // write_recipe_line's curr_letter / count (CD:46-58)
    std::string letter = "a";
    long long count = 0;
    std::string rename(const std::string& lna) {
        const size_t f = lna.find('_');
        // Python: cut = lna.find('_'); lna[:cut] with cut == -1 drops the last character
        std::string prefix, head;
        if (f == std::string::npos) {
            prefix = lna.empty() ? std::string() : lna.substr(0, lna.size() - 1);
            head = std::string();                  // lna[:0]
        } else {
            prefix = lna.substr(0, f);
            head = lna.substr(0, f + 1);
        }
        if (prefix == letter) count += 1;
        else { count = 1; letter = prefix; }
        return head + std::to_string(count);
    }
};

// Python-2 str(float): '%.12g', plus '.0' when the text has neither '.', 'e' nor 'n' (inf / nan)
std::string fstr(double x) {
    char buf[64];
    snprintf(buf, sizeof(buf), "%.12g", x);
    if (!strpbrk(buf, ".en")) strcat(buf, ".0");
    return buf;
}

inline bool is_digit(unsigned char c) { return c >= '0' && c <= '9'; }
inline bool is_space(unsigned char c) { return c == ' ' || (c >= '\t' && c <= '\r') || (c >= 0x1c && c <= 0x1f); }

// the whole text is digits, one character, digits (re.fullmatch(r'\d+.\d+')): what the reference's pattern would
// take back whole from its own output
bool time_text_whole(const std::string& s) {
    const size_t n = s.size();
    if (n < 3 || !is_digit((unsigned char)s[0]) || !is_digit((unsigned char)s[n - 1])) return false;
    size_t lead = 0;
    while (lead < n && is_digit((unsigned char)s[lead])) ++lead;
    if (lead == n) return true;                    // all digits: \d+ backs off by two characters
    // s[lead] is the one non-digit: everything behind it must be digits
    for (size_t k = lead + 1; k < n; ++k)
        if (!is_digit((unsigned char)s[k])) return false;
    return lead + 1 < n;
}

}  // namespace

struct spkdiar_replay {
    double rate = 0.0;
    std::vector<Line> lines;                       // the parsed input recipe
    std::vector<int32_t> chain_line;               // line that opens chain k
    bool single_wav = true;
    // stage B
    std::vector<Line> turns;                       // the segmentation recipe as the clustering stage reads it back
    std::string seg_text;
    long long windows = 0;
    // stage D
    std::string clu_text;
    long long speakers = 0, merges = 0;
    std::string err;
};

static int fail(spkdiar_replay* r, int code, const char* fmt, ...) {
    if (r) {
        char buf[256];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof(buf), fmt, ap);
        va_end(ap);
        r->err = buf;
    }
    return code;
}

// re.search(key + r'(\S+)'): leftmost occurrence of the key that is followed by a non-space character
static bool find_token(const char* s, size_t n, const char* key, std::string& out) {
    const size_t kl = strlen(key);
    for (size_t p = 0; p + kl < n; ++p) {
        if (memcmp(s + p, key, kl) != 0) continue;
        size_t q = p + kl;
        if (is_space((unsigned char)s[q])) continue;
        size_t e = q;
        while (e < n && !is_space((unsigned char)s[e])) ++e;
        out.assign(s + q, e - q);
        return true;
    }
    return false;
}

// re.search(key + r'(\d+.\d+)') with the UNESCAPED dot of the reference: digits, any one character but a newline,
// digits, greedy with backtracking.  The caller decides whether float() of the group is something this unit
// reproduces.
static bool find_time(const char* s, size_t n, const char* key, std::string& out) {
    const size_t kl = strlen(key);
    for (size_t p = 0; p + kl < n; ++p) {
        if (memcmp(s + p, key, kl) != 0) continue;
        const size_t q = p + kl;
        size_t run = 0;
        while (q + run < n && is_digit((unsigned char)s[q + run])) ++run;
        for (size_t k = run; k >= 1; --k) {        // \d+ gives characters back one at a time
            const size_t any = q + k;              // the '.' of the pattern
            if (any + 1 < n && s[any] != '\n' && is_digit((unsigned char)s[any + 1])) {
                size_t e = any + 1;
                while (e < n && is_digit((unsigned char)s[e])) ++e;
                out.assign(s + q, e - q);
                return true;
            }
        }
    }
    return false;
}

// float() of a matched group, for the texts this unit takes: digits with at most one '.' in between
static bool plain_number(const std::string& g, double& v) {
    int dots = 0;
    for (char c : g) {
        if (c == '.') ++dots;
        else if (!is_digit((unsigned char)c)) return false;
    }
    if (dots > 1) return false;
    v = strtod(g.c_str(), nullptr);
    return true;
}

extern "C" int spkdiar_replay_create(double rate, const char* recipe_text, int64_t len, spkdiar_replay** out) {
    if (!out) return SPKDIAR_E_ARG;
    *out = nullptr;
    if (!(rate > 0.0) || (!recipe_text && len != 0) || len < 0) return SPKDIAR_E_ARG;
    spkdiar_replay* r = new (std::nothrow) spkdiar_replay();
    if (!r) return SPKDIAR_E_NOMEM;
    r->rate = rate;
    *out = r;
    for (int64_t k = 0; k < len; ++k) {
        const unsigned char c = (unsigned char)recipe_text[k];
        if (c >= 0x80 || c == 0 || c == '\r' || c == '\v' || c == '\f' || (c >= 0x1c && c <= 0x1f))
            return fail(r, SPKDIAR_E_UNSUPPORTED, "recipe text with byte 0x%02x: line splitting / \\S / \\d follow Python's str rules there", c);
    }
    const char* s = recipe_text;
    int64_t p = 0;
    while (p < len) {
        int64_t e = p;
        while (e < len && s[e] != '\n') ++e;
        const size_t n = (size_t)(e - p);
        Line l;
        std::string t0, t1;
        // the reference evaluates audio, lna, start-time, end-time and skips the line at the first miss
        if (find_token(s + p, n, "audio=", l.audio) && find_token(s + p, n, "lna=", l.lna)
            && find_time(s + p, n, "start-time=", t0) && find_time(s + p, n, "end-time=", t1)) {
            if (!plain_number(t0, l.start) || !plain_number(t1, l.end))
                return fail(r, SPKDIAR_E_UNSUPPORTED, "time field '%s' / '%s': float() of it is left to Python", t0.c_str(), t1.c_str());
            r->lines.push_back(l);
        }
        p = e + 1;
    }
    std::string this_lna;                          // CD:362, 370-372 (never reset)
    for (size_t k = 0; k < r->lines.size(); ++k) {
        if (r->lines[k].audio != r->lines[0].audio) r->single_wav = false;
        if (r->lines[k].lna != this_lna) {
            this_lna = r->lines[k].lna;
            r->chain_line.push_back((int32_t)k);
        }
    }
    return SPKDIAR_OK;
}

extern "C" int spkdiar_replay_free(spkdiar_replay* r) {
    delete r;
    return SPKDIAR_OK;
}

extern "C" const char* spkdiar_replay_error(const spkdiar_replay* r) { return r ? r->err.c_str() : "null handle"; }

extern "C" int spkdiar_replay_info(const spkdiar_replay* r, int64_t* out6) {
    if (!r || !out6) return SPKDIAR_E_ARG;
    out6[0] = (int64_t)r->lines.size();
    out6[1] = (int64_t)r->chain_line.size();
    out6[2] = r->single_wav ? 1 : 0;
    out6[3] = (int64_t)r->turns.size();
    out6[4] = r->speakers;
    out6[5] = r->windows;
    return SPKDIAR_OK;
}

static inline int64_t clamp_frame(double v, int64_t n) {
    // int(v) truncates towards zero; min(max(., 0), n)
    if (!(v == v)) return 0;
    if (v >= 9.0e18) return n;
    if (v <= -9.0e18) return 0;
    const int64_t k = (int64_t)v;
    return k < 0 ? 0 : (k > n ? n : k);
}

extern "C" int spkdiar_replay_chains(const spkdiar_replay* r, int64_t nframes, int64_t base, int64_t* seg_a,
                                     int64_t* seg_b, int64_t cap) {
    if (!r || nframes < 0 || (cap > 0 && (!seg_a || !seg_b))) return SPKDIAR_E_ARG;
    if ((int64_t)r->chain_line.size() > cap) return SPKDIAR_E_CAPACITY;
    for (size_t k = 0; k < r->chain_line.size(); ++k) {
        const Line& l = r->lines[r->chain_line[k]];
        const int64_t a = clamp_frame(l.start * r->rate, nframes);      // CD:373: feas[int(s * rate):int(e * rate)]
        const int64_t b = clamp_frame(l.end * r->rate, nframes);
        seg_a[k] = base + a;
        seg_b[k] = base + std::max(a, b);
    }
    return SPKDIAR_OK;
}

static void write_line(std::string& text, const std::string& audio, const std::string& lna, const std::string& t0,
                       const std::string& t1, const char* tag, long long number) {
    text += "audio=";
    text += audio;
    text += " lna=";
    text += lna;
    text += " start-time=";
    text += t0;
    text += " end-time=";
    text += t1;
    text += " speaker=";
    text += tag;
    if (number >= 0) text += std::to_string(number);
    text += '\n';
}

extern "C" int spkdiar_replay_segment(spkdiar_replay* r, const spkdiar_gw_window* win, const int64_t* win_first) {
    if (!r || !win_first || (!win && win_first[r->chain_line.size()] != win_first[0])) return SPKDIAR_E_ARG;
    if (!r->single_wav) return fail(r, SPKDIAR_E_UNSUPPORTED, "a recipe naming several wavs goes through the Python replay");
    r->turns.clear();
    r->seg_text.clear();
    r->windows = 0;
    LnaState ren;
    const double rate = r->rate;
    auto emit = [&](const Line& l, double start, double end) -> bool {
        // write_recipe_line (CD:46-69): time = frames / rate + the line's start
        const std::string lna = ren.rename(l.lna);
        const std::string t0 = fstr(start / rate + l.start), t1 = fstr(end / rate + l.start);
        write_line(r->seg_text, l.audio, lna, t0, t1, "spk_turn", -1);
        // the clustering stage parses this text again: what would its four searches return?
        if (!time_text_whole(t0) || !time_text_whole(t1)) return false;
        if (l.audio.find("lna=") != std::string::npos || l.audio.find("start-time=") != std::string::npos
            || l.audio.find("end-time=") != std::string::npos || lna.find("start-time=") != std::string::npos
            || lna.find("end-time=") != std::string::npos)
            return false;
        Line t;
        t.audio = l.audio;
        t.lna = lna;
        t.start = strtod(t0.c_str(), nullptr);
        t.end = strtod(t1.c_str(), nullptr);
        r->turns.push_back(t);
        return true;
    };
    for (size_t k = 0; k < r->chain_line.size(); ++k) {
        const Line& l = r->lines[r->chain_line[k]];
        double start = 0.0;
        for (int64_t i = win_first[k]; i < win_first[k + 1]; ++i) {
            const spkdiar_gw_window& w = win[i];
            r->windows += 1;
            if (w.positive) {                      // CD:252-262
                start = w.start;
                if (!emit(l, start, start + w.maxi_fine))
                    return fail(r, SPKDIAR_E_UNSUPPORTED, "a line of the segmentation recipe would not parse back field by field (time text, or a key inside a name)");
                start += w.maxi_fine;
            }
        }
        if (!emit(l, start, (l.end - l.start) * rate))                 // CD:286-288
            return fail(r, SPKDIAR_E_UNSUPPORTED, "a line of the segmentation recipe would not parse back field by field (time text, or a key inside a name)");
    }
    return SPKDIAR_OK;
}

extern "C" int spkdiar_replay_turns(const spkdiar_replay* r, int64_t nframes, int64_t base, int64_t* seg_a,
                                    int64_t* seg_b, int64_t cap) {
    if (!r || nframes < 0 || (cap > 0 && (!seg_a || !seg_b))) return SPKDIAR_E_ARG;
    if ((int64_t)r->turns.size() > cap) return SPKDIAR_E_CAPACITY;
    for (size_t k = 0; k < r->turns.size(); ++k) {
        // CL1:282-283 + CL1:47: features[int(start * rate):int(end * rate)]
        const int64_t a = clamp_frame(r->turns[k].start * r->rate, nframes);
        const int64_t b = clamp_frame(r->turns[k].end * r->rate, nframes);
        seg_a[k] = base + a;
        seg_b[k] = base + (b > a ? b : a);
    }
    return SPKDIAR_OK;
}

extern "C" int spkdiar_replay_cluster(spkdiar_replay* r, const spkdiar_merge* merges, int64_t nmerges) {
    if (!r || nmerges < 0 || (nmerges > 0 && !merges)) return SPKDIAR_E_ARG;
    const size_t n = r->turns.size();
    if (n == 0) return fail(r, SPKDIAR_E_UNSUPPORTED, "no turn to cluster");
    std::vector<std::vector<int32_t>> spk(n);
    for (size_t k = 0; k < n; ++k) spk[k].push_back((int32_t)k);
    for (int64_t m = 0; m < nmerges; ++m) {        // CL1:216-220: extend + pop, indices in the compacted list
        const int64_t a = merges[m].a, b = merges[m].b;
        if (a < 0 || b <= a || b >= (int64_t)spk.size()) return fail(r, SPKDIAR_E_ARG, "merge %lld = (%lld, %lld) of %zu clusters", (long long)m, (long long)a, (long long)b, spk.size());
        spk[a].insert(spk[a].end(), spk[b].begin(), spk[b].end());
        spk.erase(spk.begin() + b);
    }
    const double rate = r->rate;
    struct Turn { double s, e; int32_t l, spk; };
    std::vector<Turn> all;
    all.reserve(n);
    for (size_t s = 0; s < spk.size(); ++s)
        for (int32_t l : spk[s]) all.push_back(Turn{r->turns[l].start * rate, r->turns[l].end * rate, l, (int32_t)s});
    // CL1:243-260 writes the globally smallest (start, end, line) tuple first
    std::sort(all.begin(), all.end(), [](const Turn& x, const Turn& y) {
        if (x.s != y.s) return x.s < y.s;
        if (x.e != y.e) return x.e < y.e;
        return x.l < y.l;
    });
    r->clu_text.clear();
    LnaState ren;
    for (const Turn& t : all) {
        const Line& l = r->turns[t.l];
        write_line(r->clu_text, l.audio, ren.rename(l.lna), fstr(t.s / rate + 0.0), fstr(t.e / rate + 0.0), "speaker_",
                   (long long)t.spk + 1);
    }
    r->speakers = (long long)spk.size();
    r->merges = nmerges;
    return SPKDIAR_OK;
}

extern "C" const char* spkdiar_replay_text(const spkdiar_replay* r, int32_t which, int64_t* len) {
    if (!r) return nullptr;
    const std::string& s = which == 0 ? r->seg_text : r->clu_text;
    if (len) *len = (int64_t)s.size();
    return s.c_str();
}
