// Renderer.h — presentation interface, identical to the reference's src/Renderer/Renderer.h:5-10.
#pragma once

#include <cstdint>
#include <string>

class Renderer {
public:
    virtual void renderFrame() = 0;
    virtual bool shouldStopRendering() = 0;
    virtual void send(const std::string &data) = 0;
    virtual ~Renderer() = default;
};
