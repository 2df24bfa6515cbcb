// arv2_cli -- command-line driver mirroring the reference's entry point
// (OR/main.cpp:720-777):  arv2_cli <config.json> [main|export] [export_path] [asset_dir]
// It follows the call sequence of screen() / export_audio() (OR/main.cpp:411-436,653-718):
// load config -> loadOBJ -> receiver halves -> AudioRenderer -> setters -> render ->
// convoluteAudioFile -> (export) Result.wav.  No window, no audio device (out of scope).
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <string>
#include <vector>

#include "audio_renderer.hpp"

using namespace arv2host;

int main(int argc, char** argv)
{
    if (argc < 2) { std::fprintf(stderr, "usage: %s <config.json> [main|export] [export_path] [asset_dir]\n", argv[0]); return 1; }
    const std::string mode = argc > 2 ? argv[2] : "main";
    const std::string export_path = argc > 3 ? argv[3] : "Result.wav";
    const std::string assets = argc > 4 ? argv[4] : "../../assets/models";   // hard-coded in OR/Context.cpp:190-191
    try {
        arv2_config cfg;
        check(arv2_config_load(argv[1], &cfg), "loadContext");
        std::unique_ptr<OptixModel> scene(loadOBJ(cfg.scene_file_path));
        Sphere sphere(assets + "/leftHalf.obj", assets + "/rightHalf.obj");
        float* audio = nullptr; size_t n = 0; int32_t fs = 44100, ch = 0;
        if (cfg.audio_file_path[0]) check(arv2_wav_read(cfg.audio_file_path, &audio, &n, &fs, &ch), "AudioFile::load");
        std::vector<Material> mats;
        for (int i = 0; i < cfg.n_materials; ++i) mats.push_back({cfg.material_names[i], cfg.material_absorption[i]});
        AudioRenderer renderer(scene.get(), &sphere, cfg.ir_length_in_seconds, fs, mats, {cfg.rays[0], cfg.rays[1], cfg.rays[2]});
        renderer.set_write_ir_to_file_flag(cfg.write_first_ir_to_file != 0);
        renderer.set_seed(cfg.seed);
        // OR/main.cpp:411-418
        renderer.setMonoOutput(cfg.mono != 0);
        renderer.setBasePower(cfg.base_power);
        renderer.setThresholds(cfg.ray_energy_threshold, cfg.ray_max_bounces);
        renderer.set_hrtf_absorption_rate(cfg.hrtf_absorption_rate);
        renderer.setEmitterPosInOptix({cfg.initial_emitter_pos[0], cfg.initial_emitter_pos[1], cfg.initial_emitter_pos[2]});
        renderer.setSphereCenterInOptix({cfg.initial_receiver_pos[0], cfg.initial_receiver_pos[1], cfg.initial_receiver_pos[2]}, 0.f);
        double ms = 0;
        renderer.render(&ms);
        std::printf("Time taken by trace: %g ms (%lld segments)\n", ms, renderer.last_segments());
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
