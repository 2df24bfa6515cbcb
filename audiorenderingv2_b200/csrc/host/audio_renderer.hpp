// audio_renderer.hpp -- C++ host-side mirror of the reference's renderer over the C ABI
// (include/arv2.h).  Same method names, argument meaning and error behaviour as
// class AudioRenderer (OR/AudioRenderer.h:16-152): failures throw std::runtime_error like
// CUDA_CHECK does (OR/optix7.h:8-19) instead of exit()ing.  Header-only; link libarv2.so.
#pragma once

#include <cstddef>
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
    }
    ~AudioRenderer() { arv2_destroy(ctx_); }
    AudioRenderer(const AudioRenderer&) = delete;
    AudioRenderer& operator=(const AudioRenderer&) = delete;

    void render(double* render_time = nullptr) { check(arv2_render(ctx_, render_time), "render"); flush_ir_dump(); }
    void rerender(double* render_time = nullptr) { check(arv2_rerender(ctx_, render_time), "rerender"); }

    // h_inputBufferSize is in BYTES, as in the reference (OR/AudioRenderer.cpp:663-750)
    void convoluteAudioFile(float* h_inputBuffer, size_t h_inputBufferSize, float* h_outputBuffer_left, float* h_outputBuffer_right,
                            double* convolute_time = nullptr, double* convolute_process_time = nullptr, bool reference_semantics = true)
    {
        check(arv2_convolve_file(ctx_, h_inputBuffer, h_inputBufferSize / sizeof(float), h_outputBuffer_left, h_outputBuffer_right,
                                 reference_semantics ? ARV2_CONV_REFERENCE : ARV2_CONV_LINEAR, convolute_time, convolute_process_time),
              "convoluteAudioFile");
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
    void set_write_ir_to_file_flag(bool v) { write_ir_ = v; }

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

private:
    void flush_ir_dump()
    {
        if (!write_ir_) return;   // OR/AudioRenderer.cpp:525-567
        check(arv2_write_ir_text(ctx_, "output_ir_left.txt", "output_ir_right.txt"), "write_ir_to_file");
        write_ir_ = false;
    }
    arv2_ctx* ctx_ = nullptr;
    bool write_ir_ = false;
};

} // namespace arv2host
