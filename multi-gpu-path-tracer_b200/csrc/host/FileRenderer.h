// FileRenderer.h — headless presenter: writes the frame as P6 (the README's out.ppm, reference README.md:52-58; the
// reference has no writer) and stops after `framesToRender` frames.  Stands where RemoteRenderer / LocalRenderer stand
// in the reference's main loop (src/main.cu:66-89), behind the same Renderer interface.
#pragma once

#include "../../../include/ptcore.h"
#include "Framebuffer.h"
#include "Renderer.h"
#include "RendererConfig.h"

#include <iostream>
#include <memory>

class FileRenderer : public Renderer {
public:
    FileRenderer(RendererConfig &config, std::shared_ptr<Framebuffer> &framebuffer) : config_{config}, framebuffer_{framebuffer} {}

    void renderFrame() override {
        frames_++;
        if (frames_ >= config_.framesToRender && !config_.outputPath.empty()) {
            Resolution r = framebuffer_->getResolution();
            int rc = pt_write_ppm(config_.outputPath.c_str(), framebuffer_->getRGBPtr(), r.width, r.height);
            if (rc != 0) std::cerr << "cannot write " << config_.outputPath << std::endl;
        }
    }
    bool shouldStopRendering() override { return frames_ >= config_.framesToRender; }
    void send(const std::string &data) override { std::cout << data << std::endl; }

private:
    RendererConfig &config_;
    std::shared_ptr<Framebuffer> &framebuffer_;
    unsigned int frames_ = 0;
};
