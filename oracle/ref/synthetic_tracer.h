// synthetic_tracer.h — CPU ORACLE (test infrastructure): a deterministic stand-in for PathTrace, used to drive the
// per-pixel wrapper (RayTraceCommon, the AOV writers, GetBlueNoise, the accumulation rules) with the same inputs on
// both sides of tests/test_cpu_oracle.py::test_frame_wrapper_equals_reference_text: the reference text compiled in
// ref_frame.cpp and the oracle's render_frame. It exercises what the wrapper has to handle: a varying number of
// rand() draws before the jittered-buffer coin, the blue-noise lookup, NaN and negative samples, pixels that write all
// AOVs, some and none, and the selected-pixel statistics. Sink supplies rand(), blue_noise(float[8]) and the writers.
#pragma once
#include <cstdint>
#include <cstring>

template <class Sink>
inline void synthetic_path(Sink& s, float px, float py, uint32_t frame, float out[4]) {
    uint32_t h = (uint32_t)px * 73856093u ^ (uint32_t)py * 19349663u ^ frame * 83492791u;
    h ^= h >> 15; h *= 0x2c1b3c6du; h ^= h >> 12; h *= 0x297a2d39u; h ^= h >> 15;
    float bn[8];
    s.blue_noise(bn); // GetBlueNoise(): 8 rand() draws when blue noise is off, two texture fetches + Halton when on
    float acc = 0.0f;
    for (uint32_t k = 0; k < (h & 3u); k++) acc += s.rand();
    auto unit = [](uint32_t v) { return (float)(v & 1023u) / 1023.0f; };
    float c[4] = {unit(h) * 3.0f + bn[0] + acc, unit(h >> 10) * 2.0f + bn[3] * bn[5], unit(h >> 20) + bn[6], 1.0f};
    if ((h >> 3) % 11u == 0) c[2] = -c[2];
    uint32_t nanBits = 0x7fc00000u; float nan; memcpy(&nan, &nanBits, 4);
    if ((h >> 5) % 29u == 0) c[0] = nan;     // a NaN sample is dropped whole (RayGenCommon.h:704-707)
    if ((h >> 7) % 31u == 0) c[3] = nan;
    const uint32_t mode = (h >> 9) & 3u;     // 0: escaped path, no AOV call at all
    if (mode != 0) {
        const float n3[3] = {unit(h >> 2) - 0.5f, unit(h >> 6) - 0.5f, unit(h >> 14)};
        const float p3[3] = {px * 0.01f + bn[1], py * 0.02f - bn[2], unit(h >> 18) * 50.0f};
        s.albedo(c, 1.0f);
        s.normal(n3);
        s.world_position(p3, unit(h >> 4) * 0.25f);
        if (mode >= 2) s.world_position(n3, 0.5f); // the wrapper accumulates these
        s.distance(unit(h >> 8) * 120.0f);
        s.material((int)((h >> 13) & 7u));
        if (mode == 3) s.emissive(p3);
    }
    memcpy(out, c, sizeof(c));
}
