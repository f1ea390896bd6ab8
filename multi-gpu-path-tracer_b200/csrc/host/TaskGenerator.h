// TaskGenerator.h — image -> RenderTask rectangles.
// generateEqualTasks keeps the reference's results (src/Scheduling/TaskGenerator.h:46-55,58-80: equal columns, or
// equal cells laid out by `taskLayout` with the last column/row absorbing the remainder); generateTiles is the
// fine-grained grid the DYNAMIC scheduler hands out one claim at a time.
#pragma once

#include "DevicePathTracer.h"

#include <vector>

class TaskGenerator {
public:
    TaskGenerator(int width, int height) : width_{width}, height_{height} {}
    TaskGenerator() {}

    std::vector<RenderTask> generateEqualTasks(int task_count, int width, int height) {
        std::vector<RenderTask> tasks;
        int task_width = width / task_count;
        for (int i = 0; i < task_count - 1; i++) tasks.push_back({task_width, height, i * task_width, 0});
        tasks.push_back({width - (task_count - 1) * task_width, height, (task_count - 1) * task_width, 0});
        return tasks;
    }

    std::vector<RenderTask> generateEqualTasks(int taskCount, std::vector<std::vector<int>> &taskLayout, int width, int height) {
        std::vector<RenderTask> tasks((size_t)taskCount, RenderTask{0, 0, 0, 0});
        const int rows = (int)taskLayout.size();
        const int cellH = height / rows;
        for (int r = 0; r < rows; r++) {
            const int cols = (int)taskLayout[(size_t)r].size();
            const int cellW = width / cols;
            for (int c = 0; c < cols; c++) {
                RenderTask &t = tasks[(size_t)taskLayout[(size_t)r][(size_t)c]];
                t.offset_x = cellW * c;
                t.offset_y = cellH * r;
                t.width = (c == cols - 1) ? width - t.offset_x : cellW;
                t.height = (r == rows - 1) ? height - t.offset_y : cellH;
            }
        }
        return tasks;
    }

    std::vector<RenderTask> generateTiles(int tileW, int tileH, int width, int height) {
        std::vector<RenderTask> tiles;
        for (int y = 0; y < height; y += tileH)
            for (int x = 0; x < width; x += tileW) tiles.push_back({std::min(tileW, width - x), std::min(tileH, height - y), x, y});
        return tiles;
    }

    void setRes(int width, int height) {
        width_ = width;
        height_ = height;
    }

private:
    int width_ = 0;
    int height_ = 0;
};
