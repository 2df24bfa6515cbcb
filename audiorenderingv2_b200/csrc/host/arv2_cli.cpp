// arv2_cli -- command-line driver mirroring the reference's entry point
// (OR/main.cpp:720-777):  arv2_cli <config.json> [main|export] [export_path] [asset_dir] [--gpus N]
// It follows the call sequence of screen() / export_audio() (OR/main.cpp:411-436,653-718):
// load config -> loadOBJ -> receiver halves -> AudioRenderer -> setters -> render ->
// convoluteAudioFile -> (export) Result.wav.  No window, no audio device (out of scope).
// --gpus N: the render is sharded over devices 0..N-1 of this box (one process, one host thread and one NCCL rank
// per device, arv2_multi_*); everything after the render runs on device 0, which holds the full IR.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "audio_renderer.hpp"

using namespace arv2host;

template <class R>
static void configure(R& renderer, const arv2_config& cfg)
{
    renderer.set_seed(cfg.seed);
    // OR/main.cpp:411-418
    renderer.setMonoOutput(cfg.mono != 0);
    renderer.setBasePower(cfg.base_power);
    renderer.setThresholds(cfg.ray_energy_threshold, cfg.ray_max_bounces);
    renderer.set_hrtf_absorption_rate(cfg.hrtf_absorption_rate);
    renderer.setEmitterPosInOptix({cfg.initial_emitter_pos[0], cfg.initial_emitter_pos[1], cfg.initial_emitter_pos[2]});
    renderer.setSphereCenterInOptix({cfg.initial_receiver_pos[0], cfg.initial_receiver_pos[1], cfg.initial_receiver_pos[2]}, 0.f);
}

int main(int argc, char** argv)
{
    int gpus = 1;
    std::vector<std::string> pos;
    for (int i = 1; i < argc; ++i) {
        if (!std::strcmp(argv[i], "--gpus") && i + 1 < argc) gpus = std::atoi(argv[++i]);
        else pos.push_back(argv[i]);
    }
    if (pos.empty() || gpus < 1) { std::fprintf(stderr, "usage: %s <config.json> [main|export] [export_path] [asset_dir] [--gpus N]\n", argv[0]); return 1; }
    const std::string mode = pos.size() > 1 ? pos[1] : "main";
    const std::string export_path = pos.size() > 2 ? pos[2] : "Result.wav";
    const std::string assets = pos.size() > 3 ? pos[3] : "../../assets/models";   // hard-coded in OR/Context.cpp:190-191
    try {
        arv2_config cfg;
        check(arv2_config_load(pos[0].c_str(), &cfg), "loadContext");
        std::unique_ptr<OptixModel> scene(loadOBJ(cfg.scene_file_path));
        Sphere sphere(assets + "/leftHalf.obj", assets + "/rightHalf.obj");
        float* audio = nullptr; size_t n = 0; int32_t fs = 44100, ch = 0;
        if (cfg.audio_file_path[0]) check(arv2_wav_read(cfg.audio_file_path, &audio, &n, &fs, &ch), "AudioFile::load");
        std::vector<Material> mats;
        for (int i = 0; i < cfg.n_materials; ++i) mats.push_back({cfg.material_names[i], cfg.material_absorption[i]});
        const Vec3 rays{cfg.rays[0], cfg.rays[1], cfg.rays[2]};
        std::unique_ptr<AudioRenderer> single;
        std::unique_ptr<AudioRendererMulti> multi;
        double ms = 0;
        long long segments = 0;
        if (gpus == 1) {
            single.reset(new AudioRenderer(scene.get(), &sphere, cfg.ir_length_in_seconds, fs, mats, rays));
            configure(*single, cfg);
            single->set_write_ir_to_file_flag(cfg.write_first_ir_to_file != 0);      // OR/Context.cpp:229
            single->render(&ms);
            segments = single->last_segments();
        } else {
            std::vector<int> devices;
            for (int d = 0; d < gpus; ++d) devices.push_back(d);
            multi.reset(new AudioRendererMulti(scene.get(), &sphere, cfg.ir_length_in_seconds, fs, mats, rays, devices));
            configure(*multi, cfg);
            multi->render(&ms);
            segments = multi->last_segments();
            if (cfg.write_first_ir_to_file) check(arv2_write_ir_text(multi->device(0).handle(), "output_ir_left.txt", "output_ir_right.txt"), "write_ir_to_file");
        }
        AudioRenderer& renderer = gpus == 1 ? *single : multi->device(0);
        renderer.set_write_output_to_file_flag(cfg.write_first_output_to_file != 0);  // OR/Context.cpp:230
        std::printf("Time taken by trace: %g ms (%lld segments, %d GPU%s)\n", ms, segments, gpus, gpus > 1 ? "s" : "");
        if (audio) {
            std::vector<float> l(n), r(n);
            double t = 0, tp = 0;
            renderer.convoluteAudioFile(audio, n * sizeof(float), l.data(), r.data(), &t, &tp);
            std::printf("Time taken just to convolute: %g ms\nTime taken for convolution process: %g ms\n", t, tp);
            if (mode == "export") {
                check(arv2_wav_write_stereo_normalized(export_path.c_str(), l.data(), r.data(), n, fs), "export_audio");
                std::printf("wrote %s\n", export_path.c_str());
            }
            arv2_free(audio);
        }
    } catch (const std::exception& e) {
        std::fprintf(stderr, "Exception caught: %s\n", e.what());      // OR/main.cpp:771-775
        return 1;
    }
    return 0;
}
