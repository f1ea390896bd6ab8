// RenderTask.h — one rectangle of the frame handed to one (GPU, stream) worker; the reference declares it at the top of
// src/DevicePathTracer.h:19-25 (same members, same order: aggregate initialisation {width, height, offset_x, offset_y} is
// used all over its RenderManager).  offset_y counts from the bottom row of the image, as the kernels do; `time` is the
// worker's render time of the previous frame in milliseconds, the input of the DSFL / DSDL schedulers.
#pragma once

struct RenderTask {
    int width;
    int height;
    int offset_x;
    int offset_y;
    int time = 0;
};
