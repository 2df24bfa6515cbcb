// wav.cpp -- minimal RIFF/WAVE I/O for libarv2.
// Decode rule of the reference's AudioFile<float>::load as used by Context.cpp:198-213
// (channel 0 only): 16-bit PCM -> sample / 32768 (AudioFile.h:1242-1245), IEEE float32
// passed through (AudioFile.h:617-618).  Writer = export mode's Result.wav
// (OR/main.cpp:628-718): per-channel min-max normalisation to [-1, 1], 16-bit stereo.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>

#include "../arv2_internal.h"

namespace arv2 {

int wav_read(const std::string& path, float** samples, size_t* n, int32_t* rate, int32_t* channels, std::string* err)
{
    std::ifstream in(path, std::ios::binary);
    if (!in) { *err = "cannot open " + path; return ARV2_ERR_IO; }
    std::vector<unsigned char> d((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
    if (d.size() < 12 || std::memcmp(d.data(), "RIFF", 4) || std::memcmp(d.data() + 8, "WAVE", 4)) { *err = "not a RIFF/WAVE file"; return ARV2_ERR_IO; }
    auto u16 = [&](size_t o) { return (unsigned)(d[o] | (d[o + 1] << 8)); };
    auto u32 = [&](size_t o) { return (unsigned)(d[o] | (d[o + 1] << 8) | (d[o + 2] << 16) | ((unsigned)d[o + 3] << 24)); };
    size_t pos = 12, data_off = 0, data_len = 0;
    unsigned tag = 0, ch = 0, sr = 0, bits = 0;
    while (pos + 8 <= d.size()) {
        const unsigned size = u32(pos + 4);
        if (!std::memcmp(d.data() + pos, "fmt ", 4) && pos + 8 + 16 <= d.size()) {
            tag = u16(pos + 8); ch = u16(pos + 10); sr = u32(pos + 12); bits = u16(pos + 22);
        } else if (!std::memcmp(d.data() + pos, "data", 4)) {
            data_off = pos + 8; data_len = std::min<size_t>(size, d.size() - data_off);
        }
        pos += 8 + (size_t)size + (size & 1u);
    }
    if (!ch || !data_off) { *err = "missing fmt/data chunk"; return ARV2_ERR_IO; }
    const size_t bps = bits / 8;
    if (!((tag == 1 && bits == 16) || (tag == 3 && bits == 32))) { *err = "unsupported WAV encoding"; return ARV2_ERR_IO; }
    const size_t frames = data_len / (bps * ch);
    float* out = (float*)std::malloc(std::max<size_t>(1, frames) * sizeof(float));
    if (!out) { *err = "out of memory"; return ARV2_ERR_NOMEM; }
    for (size_t f = 0; f < frames; ++f) {
        const size_t o = data_off + f * bps * ch;           // channel 0
        if (tag == 1) { const int16_t s = (int16_t)u16(o); out[f] = (float)s / 32768.f; }
        else { std::memcpy(&out[f], d.data() + o, 4); }
    }
    *samples = out; *n = frames; *rate = (int32_t)sr; *channels = (int32_t)ch;
    return ARV2_OK;
}

int wav_write_stereo_normalized(const std::string& path, const float* l, const float* r, size_t n, int32_t rate, std::string* err)
{
    auto norm = [n](const float* a, std::vector<float>* o) {
        float lo = 0.f, hi = 0.f;
        if (n) { lo = hi = a[0]; }
        for (size_t i = 1; i < n; ++i) { lo = std::min(lo, a[i]); hi = std::max(hi, a[i]); }
        o->resize(n);
        if (lo == hi) return false;   // normalizeToRangeMinusOneToOne throws (OR/main.cpp:641-643)
        for (size_t i = 0; i < n; ++i) (*o)[i] = 2.f * ((a[i] - lo) / (hi - lo)) - 1.f;
        return true;
    };
    std::vector<float> a, b;
    if (!norm(l, &a) || !norm(r, &b)) { *err = "Cannot normalize: all elements in the vector are the same."; return ARV2_ERR_INVALID; }
    std::ofstream out(path, std::ios::binary);
    if (!out) { *err = "cannot open " + path; return ARV2_ERR_IO; }
    const uint32_t data_bytes = (uint32_t)(n * 4);
    auto w32 = [&](uint32_t v) { out.write((const char*)&v, 4); };
    auto w16 = [&](uint16_t v) { out.write((const char*)&v, 2); };
    out.write("RIFF", 4); w32(36 + data_bytes); out.write("WAVE", 4);
    out.write("fmt ", 4); w32(16); w16(1); w16(2); w32((uint32_t)rate); w32((uint32_t)rate * 4); w16(4); w16(16);
    out.write("data", 4); w32(data_bytes);
    for (size_t i = 0; i < n; ++i) {
        auto q = [](float v) { v = v < -1.f ? -1.f : (v > 1.f ? 1.f : v); return (int16_t)(v * 32767.f); };
        w16((uint16_t)q(a[i])); w16((uint16_t)q(b[i]));
    }
    return ARV2_OK;
}

} // namespace arv2
