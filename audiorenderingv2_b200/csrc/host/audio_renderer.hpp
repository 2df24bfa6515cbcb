// audio_renderer.hpp -- C++ host-side mirror of the reference's renderer over the C ABI
// (include/arv2.h).  Same method names, argument meaning and error behaviour as
// class AudioRenderer (OR/AudioRenderer.h:16-152): failures throw std::runtime_error like
// CUDA_CHECK does (OR/optix7.h:8-19) instead of exit()ing.  Header-only; link libarv2.so.
#pragma once

#include <chrono>
#include <cstddef>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../../include/arv2.h"

namespace arv2host {

inline void check(int rc, const char* what)
{
    if (rc != ARV2_OK) throw std::runtime_error(std::string(what) + ": " + arv2_last_error());
}

struct Vec3 { float x, y, z; };

// struct Material (OR/LaunchParams.h:14-18)
struct Material { std::string name; float mat_absorption; };

// struct OptixModel (OR/OptixModel.h:21-32)
class OptixModel {
public:
    explicit OptixModel(arv2_scene* h) : h_(h) {}
    ~OptixModel() { arv2_scene_destroy(h_); }
    OptixModel(const OptixModel&) = delete;
    OptixModel& operator=(const OptixModel&) = delete;
    arv2_scene* handle() const { return h_; }
private:
    arv2_scene* h_;
};

// loadOBJ (OR/OptixModel.cpp:75-151); throws like the reference on unreadable files or
// files without materials.
inline OptixModel* loadOBJ(const std::string& objFile)
{
    arv2_scene* s = nullptr;
    check(arv2_scene_load_obj(objFile.c_str(), &s), "loadOBJ");
    return new OptixModel(s);
}

// class Sphere / HalfSphere (OR/Sphere.cpp, OR/HalfSphere.cpp)
class Sphere {
public:
    Sphere(const std::string& left_obj, const std::string& right_obj) { check(arv2_receiver_load(left_obj.c_str(), right_obj.c_str(), &h_), "HalfSphere"); }
    ~Sphere() { arv2_receiver_destroy(h_); }
    Sphere(const Sphere&) = delete;
    Sphere& operator=(const Sphere&) = delete;
    arv2_receiver* handle() const { return h_; }
private:
    arv2_receiver* h_ = nullptr;
};

// template <typename T> class CircularBuffer (OR/CircularBuffer.h), the double instance the live path uses
class CircularBuffer {
public:
    explicit CircularBuffer(size_t size) { check(arv2_ring_create(size, &h_), "CircularBuffer"); }
    ~CircularBuffer() { arv2_ring_destroy(h_); }
    CircularBuffer(const CircularBuffer&) = delete;
    CircularBuffer& operator=(const CircularBuffer&) = delete;
    void add(const double* values, size_t n) { check(arv2_ring_add(h_, values, n), "CircularBuffer::add"); }
    void get_and_reset(double* out, size_t n) { check(arv2_ring_get_and_reset(h_, out, n), "CircularBuffer::get_and_reset"); }
    arv2_ring* handle() const { return h_; }
private:
    arv2_ring* h_ = nullptr;
};

class AudioRenderer {
public:
    // AudioRenderer(const OptixModel*, unsigned ir_length_in_seconds, int sample_rate,
    //               std::vector<Material>, gdt::vec3f rays_per_dimension)  OR/AudioRenderer.cpp:60-93
    AudioRenderer(const OptixModel* model, const Sphere* sphere, unsigned ir_length_in_seconds, int sample_rate,
                  const std::vector<Material>& materials, Vec3 rays_per_dimension, int device = 0, bool path_cache = false)
    {
        std::vector<arv2_material> m(materials.size());
        for (size_t i = 0; i < materials.size(); ++i) {
            m[i].name = materials[i].name.c_str();
            for (int b = 0; b < ARV2_MAX_BANDS; ++b) m[i].mat_absorption[b] = materials[i].mat_absorption;
            m[i].scattering = 0.f;
        }
        arv2_renderer_desc d{};
        d.ir_length_in_seconds = ir_length_in_seconds; d.sample_rate = sample_rate;
        d.rays_x = (int)rays_per_dimension.x; d.rays_y = (int)rays_per_dimension.y; d.rays_z = (int)rays_per_dimension.z;
        d.bands = 1; d.device = device; d.path_cache = path_cache ? 1 : 0;
        d.materials = m.data(); d.n_materials = (int)m.size();
        check(arv2_create(model->handle(), sphere ? sphere->handle() : nullptr, &d, &ctx_), "AudioRenderer");
        device_ = device;
    }
    ~AudioRenderer() { arv2_stream_close(live_); if (owned_) arv2_destroy(ctx_); }
    // a context owned elsewhere (one device of AudioRendererMulti)
    explicit AudioRenderer(arv2_ctx* borrowed) : ctx_(borrowed), owned_(false) {}
    AudioRenderer(const AudioRenderer&) = delete;
    AudioRenderer& operator=(const AudioRenderer&) = delete;

    void render(double* render_time = nullptr) { check(arv2_render(ctx_, render_time), "render"); after_render(); }
    void rerender(double* render_time = nullptr) { check(arv2_rerender(ctx_, render_time), "rerender"); }

    // h_inputBufferSize is in BYTES, as in the reference (OR/AudioRenderer.cpp:663-750)
    void convoluteAudioFile(float* h_inputBuffer, size_t h_inputBufferSize, float* h_outputBuffer_left, float* h_outputBuffer_right,
                            double* convolute_time = nullptr, double* convolute_process_time = nullptr, bool reference_semantics = true)
    {
        check(arv2_convolve_file(ctx_, h_inputBuffer, h_inputBufferSize / sizeof(float), h_outputBuffer_left, h_outputBuffer_right,
                                 reference_semantics ? ARV2_CONV_REFERENCE : ARV2_CONV_LINEAR, convolute_time, convolute_process_time),
              "convoluteAudioFile");
        if (write_output_) {      // OR/AudioRenderer.cpp:720-744: once, then the flag resets itself
            check(arv2_write_convolved_text("output_convolute_left.txt", "output_convolute_right.txt", h_outputBuffer_left, h_outputBuffer_right,
                                            h_inputBufferSize / sizeof(float)), "write_output_to_file");
            write_output_ = false;
        }
    }

    // convoluteLiveInput (OR/AudioRenderer.cpp:593-661): h_inputBufferSize is in BYTES (OR/main.cpp:112 passes
    // nBufferFrames * sizeof(double)); the convolved block goes into the ring interleaved LRLR with the reference's gain
    // of 2.  The reference convolves every callback with an FFT of the whole IR and mallocs / plans per call; here a
    // 1-source partitioned stream convolver (512-sample blocks) lives with the renderer and carries the tail itself.
    void convoluteLiveInput(double* h_inputBuffer, size_t h_inputBufferSize, CircularBuffer* h_circularOutputBuffer)
    {
        if (!live_) check(arv2_stream_open(device_, 1, kLiveBlock, ir_length(), &live_), "convoluteLiveInput");
        if (live_ir_dirty_) {
            float *dl = nullptr, *dr = nullptr;
            check(arv2_ir_device(ctx_, &dl, &dr), "convoluteLiveInput");
            check(arv2_stream_set_ir_device(live_, 0, dl, dr), "convoluteLiveInput");
            live_ir_dirty_ = false;
        }
        check(arv2_live_callback(live_, h_inputBuffer, h_inputBufferSize / sizeof(double), h_circularOutputBuffer->handle()), "convoluteLiveInput");
    }

    void setEmitterPosInOptix(Vec3 pos) { check(arv2_set_emitter(ctx_, pos.x, pos.y, pos.z), "setEmitterPosInOptix"); }
    // placeReceiver + setSphereCenterInOptix in one call (the reference always pairs them,
    // OR/AudioRenderer.cpp:793-794, OR/main.cpp:411-417)
    void setSphereCenterInOptix(Vec3 center, float camera_global_angle = 0.f) { check(arv2_set_receiver(ctx_, center.x, center.y, center.z, camera_global_angle), "setSphereCenterInOptix"); }
    void setThresholds(float energy, unsigned int max_bounces) { check(arv2_set_thresholds(ctx_, energy, max_bounces), "setThresholds"); }
    void set_hrtf_absorption_rate(float r) { check(arv2_set_hrtf_absorption_rate(ctx_, r), "set_hrtf_absorption_rate"); }
    void setBasePower(float p) { check(arv2_set_base_power(ctx_, p), "setBasePower"); }
    void setMonoOutput(bool v) { check(arv2_set_mono(ctx_, v ? 1 : 0), "setMonoOutput"); }
    void set_seed(unsigned long long s) { check(arv2_set_seed(ctx_, s), "set_seed"); }
    void set_coherent_order(bool on) { check(arv2_set_coherent_order(ctx_, on ? 1 : 0), "set_coherent_order"); }
    void set_shard_mode(int mode) { check(arv2_set_shard_mode(ctx_, mode), "set_shard_mode"); }
    void set_sweep_min_rays(int64_t n) { check(arv2_set_sweep_min_rays(ctx_, n), "set_sweep_min_rays"); }
    void set_write_ir_to_file_flag(bool v) { write_ir_ = v; }
    void set_write_output_to_file_flag(bool v) { write_output_ = v; }
    // OR/AudioRenderer.cpp:236-239,535-540: IR dumps go to experimentation/output_ir_{left,right}_<clock ticks>.txt
    void enable_experimentation() { experimentation_mode_ = true; }

    // full_render_cycle (OR/AudioRenderer.cpp:790-798)
    void full_render_cycle(std::mutex* mutex, Vec3 camera_central_point, float camera_global_angle, float* audio_samples,
                           size_t size_of_audio, float* outputBuffer_left, float* outputBuffer_right)
    {
        std::lock_guard<std::mutex> lock(*mutex);
        setSphereCenterInOptix(camera_central_point, camera_global_angle);
        render();
        convoluteAudioFile(audio_samples, size_of_audio, outputBuffer_left, outputBuffer_right);
    }

    void getIROnHostMem(float* ir_left, float* ir_right) { check(arv2_get_ir(ctx_, ir_left, ir_right), "getIROnHostMem"); }
    int ir_length() const { int n = 0, b = 0; arv2_ir_length(ctx_, &n, &b); return n; }
    long long last_segments() const { int64_t s = 0; arv2_last_segments(ctx_, &s); return (long long)s; }
    arv2_ctx* handle() const { return ctx_; }

    void ir_changed() { live_ir_dirty_ = true; }

protected:
    void after_render()
    {
        live_ir_dirty_ = true;
        if (!write_ir_) return;   // OR/AudioRenderer.cpp:525-567
        std::string l = "output_ir_left.txt", r = "output_ir_right.txt";
        if (experimentation_mode_) {
            const auto ticks = std::chrono::system_clock::now().time_since_epoch().count();
            l = "experimentation/output_ir_left_" + std::to_string(ticks) + ".txt";
            r = "experimentation/output_ir_right_" + std::to_string(ticks) + ".txt";
        }
        check(arv2_write_ir_text(ctx_, l.c_str(), r.c_str()), "write_ir_to_file");
        write_ir_ = false;
    }
    static constexpr int kLiveBlock = 512;
    arv2_ctx* ctx_ = nullptr;
    bool owned_ = true;
    int device_ = 0;
    arv2_stream* live_ = nullptr;
    bool live_ir_dirty_ = true;
    bool write_ir_ = false, write_output_ = false, experimentation_mode_ = false;
};

// The same renderer over several GPUs of one box, in one process (the reference application is one process,
// OR/main.cpp:720-777): rays shard over the devices, the IR histograms are summed with NCCL inside the library
// (arv2_multi_*).  device(i) is an AudioRenderer over device i's context: parameters set through the forwarding
// setters below reach every device; the IR, the file convolver and the live convolver are device(0)'s.
class AudioRendererMulti {
public:
    AudioRendererMulti(const OptixModel* model, const Sphere* sphere, unsigned ir_length_in_seconds, int sample_rate,
                       const std::vector<Material>& materials, Vec3 rays_per_dimension, const std::vector<int>& devices)
    {
        std::vector<arv2_material> m(materials.size());
        for (size_t i = 0; i < materials.size(); ++i) {
            m[i].name = materials[i].name.c_str();
            for (int b = 0; b < ARV2_MAX_BANDS; ++b) m[i].mat_absorption[b] = materials[i].mat_absorption;
            m[i].scattering = 0.f;
        }
        arv2_renderer_desc d{};
        d.ir_length_in_seconds = ir_length_in_seconds; d.sample_rate = sample_rate;
        d.rays_x = (int)rays_per_dimension.x; d.rays_y = (int)rays_per_dimension.y; d.rays_z = (int)rays_per_dimension.z;
        d.bands = 1;
        d.materials = m.data(); d.n_materials = (int)m.size();
        std::vector<int32_t> dev(devices.begin(), devices.end());
        check(arv2_multi_create(model->handle(), sphere ? sphere->handle() : nullptr, &d, dev.data(), (int32_t)dev.size(), &multi_), "AudioRendererMulti");
        for (int32_t i = 0; i < arv2_multi_size(multi_); ++i) dev_.emplace_back(new AudioRenderer(arv2_multi_ctx(multi_, i)));
    }
    ~AudioRendererMulti() { dev_.clear(); arv2_multi_destroy(multi_); }
    AudioRendererMulti(const AudioRendererMulti&) = delete;
    AudioRendererMulti& operator=(const AudioRendererMulti&) = delete;

    size_t size() const { return dev_.size(); }
    AudioRenderer& device(size_t i) { return *dev_[i]; }
    void render(double* render_time = nullptr) { check(arv2_multi_render(multi_, render_time), "render"); for (auto& r : dev_) r->ir_changed(); }
    void setEmitterPosInOptix(Vec3 p) { for (auto& r : dev_) r->setEmitterPosInOptix(p); }
    void setSphereCenterInOptix(Vec3 c, float angle = 0.f) { for (auto& r : dev_) r->setSphereCenterInOptix(c, angle); }
    void setThresholds(float e, unsigned int mb) { for (auto& r : dev_) r->setThresholds(e, mb); }
    void set_hrtf_absorption_rate(float v) { for (auto& r : dev_) r->set_hrtf_absorption_rate(v); }
    void setBasePower(float v) { for (auto& r : dev_) r->setBasePower(v); }
    void setMonoOutput(bool v) { for (auto& r : dev_) r->setMonoOutput(v); }
    void set_seed(unsigned long long s) { for (auto& r : dev_) r->set_seed(s); }
    long long last_segments() const { long long t = 0; for (auto& r : dev_) t += r->last_segments(); return t; }

private:
    arv2_multi* multi_ = nullptr;
    std::vector<std::unique_ptr<AudioRenderer>> dev_;
};

} // namespace arv2host
