// playback.cpp -- the callers either side of the hot path: the
// re-render trigger of the GL loop, the receiver yaw convention, the RtAudio playback
// callback's buffer contract and the live path's CircularBuffer.  Host-only C++.
#include <cmath>
#include <cstring>
#include <vector>

#include "../arv2_internal.h"

using namespace arv2;

struct arv2_rerender_policy {
    float distance_threshold, angle_threshold;
    float last_pos[3];
    float last_angle;
    bool timer_set = false;
    double last_render_time = 0.0;
};

struct arv2_ring {
    std::vector<double> buffer;
    size_t index = 0;
};

extern "C" {

// Camera::calculate_global_angle, OR/Camera.cpp:31-41
float arv2_global_angle(float ox, float oz)
{
    float a = std::atan2(oz, ox) * 57.295779513082320876798154814105f;     // glm::degrees
    if (a < 0.f) a += 360.0f;
    return a;
}

int arv2_policy_create(float distance_threshold, float angle_threshold_deg, const float start_pos[3], float start_angle_deg,
                       arv2_rerender_policy** out)
{
    if (!out || !start_pos) { set_error("arv2_policy_create: null argument"); return ARV2_ERR_INVALID; }
    auto* p = new arv2_rerender_policy;
    p->distance_threshold = distance_threshold; p->angle_threshold = angle_threshold_deg;
    std::memcpy(p->last_pos, start_pos, sizeof p->last_pos);
    p->last_angle = start_angle_deg;
    *out = p;
    return ARV2_OK;
}

// OR/main.cpp:470-498
int arv2_policy_update(arv2_rerender_policy* p, const float pos[3], float angle_deg, double now_s, int32_t is_rendering)
{
    if (!p || !pos) return 0;
    const float dx = pos[0] - p->last_pos[0], dy = pos[1] - p->last_pos[1], dz = pos[2] - p->last_pos[2];
    const float dist = std::sqrt(dx * dx + dy * dy + dz * dz);
    if (dist > 0.f && !p->timer_set) { p->last_render_time = std::floor(now_s); p->timer_set = true; }   // time(&last_render_time)
    const bool far_enough = dist > p->distance_threshold;
    float diff = std::fabs(p->last_angle - angle_deg);
    if (diff > 180.0f) diff = 360.0f - diff;
    const bool turned = diff > p->angle_threshold;
    const bool timed_out = p->timer_set && (std::floor(now_s) - p->last_render_time) > 1.0;      // difftime(time(NULL), t) > 1
    if ((far_enough || turned || timed_out) && !is_rendering) {
        p->timer_set = false;
        p->last_angle = angle_deg;
        std::memcpy(p->last_pos, pos, sizeof p->last_pos);
        return 1;
    }
    return 0;
}

void arv2_policy_destroy(arv2_rerender_policy* p) { delete p; }

// audioHandler, OR/main.cpp:69-97
int64_t arv2_playback_fill(double* out, uint32_t n_frames, double stream_time, int32_t sample_rate, const float* left,
                           const float* right, size_t n_samples, size_t output_buffer_len, float volume)
{
    if (!out || !left || !right || n_samples == 0) return 0;
    const size_t next = (size_t)((int)(stream_time * sample_rate) % (long long)n_samples);
    int64_t written = 0;
    for (unsigned i = 0; i < n_frames * 2; ++i) {
        if (i + next >= output_buffer_len) break;
        if (i + next >= n_samples) break;              // the reference reads past its buffers here
        const float v = (i % 2 == 0) ? left[i + next] : right[i + next];
        out[written++] = (double)(v * 100 * volume);
    }
    return written;
}

// CircularBuffer<double>, OR/CircularBuffer.h
int arv2_ring_create(size_t size, arv2_ring** out)
{
    if (!out || size == 0) { set_error("arv2_ring_create: bad argument"); return ARV2_ERR_INVALID; }
    auto* r = new arv2_ring;
    r->buffer.assign(size, 0.0);
    *out = r;
    return ARV2_OK;
}

int arv2_ring_add(arv2_ring* r, const double* values, size_t n)
{
    if (!r || !values) { set_error("arv2_ring_add: null argument"); return ARV2_ERR_INVALID; }
    size_t idx = r->index;
    for (size_t i = 0; i < n; ++i) { r->buffer[idx] += values[i]; idx = (idx + 1) % r->buffer.size(); }
    return ARV2_OK;                                   // index is not advanced (CircularBuffer.h:13-20)
}

int arv2_ring_get_and_reset(arv2_ring* r, double* out, size_t n)
{
    if (!r || !out) { set_error("arv2_ring_get_and_reset: null argument"); return ARV2_ERR_INVALID; }
    if (n > r->buffer.size()) { set_error("Requested more elements than present in the buffer"); return ARV2_ERR_INVALID; }
    for (size_t i = 0; i < n; ++i) {
        double& v = r->buffer[(r->index + i) % r->buffer.size()];
        out[i] = v; v = 0.0;
    }
    r->index = (r->index + n) % r->buffer.size();
    return ARV2_OK;
}

void arv2_ring_destroy(arv2_ring* r) { delete r; }

} // extern "C"
