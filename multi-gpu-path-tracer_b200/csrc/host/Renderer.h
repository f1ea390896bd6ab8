// Renderer.h — what the render loop needs from a presenter (reference: src/Renderer/Renderer.h:5-10, implemented there by
// the websocket RemoteRenderer and the GLFW LocalRenderer; here by the headless FileRenderer).  The three operations and
// their signatures are the reference's, so either of its presenters could be compiled against this header:
//   renderFrame()          present the frame RenderManager has just finished (encode + send, blit, or write a file);
//   shouldStopRendering()  polled by the main loop between frames (src/main.cu:66);
//   send(text)             side channel for text messages — the monitor thread's "RENDER_STATS#..." lines go through it.
// A virtual destructor is added: presenters are owned through this interface.
#pragma once

#include <cstdint>
#include <string>

class Renderer {
public:
    virtual ~Renderer() = default;
    virtual void renderFrame() = 0;
    virtual bool shouldStopRendering() = 0;
    virtual void send(const std::string &data) = 0;
};
