// TaskGenerator.h — image -> RenderTask rectangles.
// generateEqualTasks keeps the reference's results (src/Scheduling/TaskGenerator.h:46-55,58-80: equal columns, or
// equal cells laid out by `taskLayout` with the last column/row absorbing the remainder); generateTiles is the
// fine-grained grid the DYNAMIC scheduler hands out one claim at a time; adjustTasksDSFL / bisectTasksDSDL are the two
// time-driven schedulers of the reference (src/RenderManager.h:264-408, :546-639) as pure functions of the tasks and
// their last render times, so they can be tested without a GPU (csrc/host/task_generator_test.cpp).
#pragma once

#include "RenderTask.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
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

    // DSFL: keep the layout, move every border by at most one thread block per frame towards the position that would have
    // equalised the previous frame's task times (columns inside each row, then the row heights).
    void adjustTasksDSFL(std::vector<RenderTask> &tasks, const std::vector<std::vector<int>> &layout, int W, int H, int bx, int by) {
        for (const auto &row : layout) {
            if (row.empty()) continue;
            double total = 0;
            for (int id : row) total += std::max(1, tasks[(size_t)id].time);
            int x = 0;
            for (size_t c = 0; c + 1 < row.size(); c++) {
                RenderTask &t = tasks[(size_t)row[c]];
                // where the border should be so that this column takes total/n: scale the column by target/actual time
                double target = total / (double)row.size();
                int ideal = (int)std::lround(t.width * target / std::max(1, t.time));
                int step = std::clamp(ideal - t.width, -bx, bx);
                int remainingCols = (int)(row.size() - 1 - c);
                int nw = std::clamp(t.width + step, 1, std::max(1, W - x - remainingCols));
                t.offset_x = x;
                t.width = nw;
                x += nw;
            }
            RenderTask &last = tasks[(size_t)row.back()];
            last.offset_x = x;
            last.width = W - x;
        }
        std::vector<double> rowTime;
        double total = 0;
        for (const auto &row : layout) {
            double s = 0;
            for (int id : row) s += std::max(1, tasks[(size_t)id].time);
            rowTime.push_back(s);
            total += s;
        }
        int y = 0;
        for (size_t r = 0; r < layout.size(); r++) {
            if (layout[r].empty()) continue;
            int hgt;
            int cur = tasks[(size_t)layout[r][0]].height;
            if (r + 1 < layout.size()) {
                double target = total / (double)layout.size();
                int ideal = (int)std::lround(cur * target / std::max(1.0, rowTime[r]));
                int step = std::clamp(ideal - cur, -by, by);
                hgt = std::clamp(cur + step, 1, std::max(1, H - y - (int)(layout.size() - 1 - r)));
            } else {
                hgt = H - y;
            }
            for (int id : layout[r]) {
                tasks[(size_t)id].offset_y = y;
                tasks[(size_t)id].height = hgt;
            }
            y += hgt;
        }
    }

    // DSDL: spread each task's time over the thread blocks it covered, then bisect the block grid recursively (alternating
    // axis) at the time-weighted median until there is one rectangle per worker.
    std::vector<RenderTask> bisectTasksDSDL(const std::vector<RenderTask> &tasks, int count, int W, int H, int bx, int by) {
        const int gw = (W + bx - 1) / bx, gh = (H + by - 1) / by;
        std::vector<float> cost((size_t)gw * gh, 0.f);
        for (const RenderTask &t : tasks) {
            int x0 = t.offset_x / bx, x1 = std::min(gw, (t.offset_x + t.width + bx - 1) / bx);
            int y0 = t.offset_y / by, y1 = std::min(gh, (t.offset_y + t.height + by - 1) / by);
            int n = std::max(1, (x1 - x0) * (y1 - y0));
            for (int yy = y0; yy < y1; yy++)
                for (int xx = x0; xx < x1; xx++) cost[(size_t)yy * gw + xx] += (float)std::max(1, t.time) / (float)n;
        }
        std::vector<RenderTask> out;
        struct Rect { int x0, y0, x1, y1, count; bool vert; };
        std::vector<Rect> stack{{0, 0, gw, gh, count, true}};
        while (!stack.empty()) {
            Rect r = stack.back();
            stack.pop_back();
            if (r.count <= 1 || (r.x1 - r.x0 < 2 && r.y1 - r.y0 < 2)) {  // one worker left, or a single block that cannot be cut
                int px0 = r.x0 * bx, py0 = r.y0 * by, px1 = std::min(W, r.x1 * bx), py1 = std::min(H, r.y1 * by);
                out.push_back({px1 - px0, py1 - py0, px0, py0});
                continue;
            }
            const int left = r.count / 2;
            double total = 0;
            for (int yy = r.y0; yy < r.y1; yy++)
                for (int xx = r.x0; xx < r.x1; xx++) total += cost[(size_t)yy * gw + xx];
            const double target = total * (double)left / (double)r.count;
            bool vert = r.vert;
            if (vert && r.y1 - r.y0 < 2) vert = false;
            if (!vert && r.x1 - r.x0 < 2) vert = true;
            double acc = 0;
            if (vert) {
                int cut = r.y0 + 1;
                for (int yy = r.y0; yy < r.y1 - 1; yy++) {
                    for (int xx = r.x0; xx < r.x1; xx++) acc += cost[(size_t)yy * gw + xx];
                    cut = yy + 1;
                    if (acc >= target) break;
                }
                stack.push_back({r.x0, cut, r.x1, r.y1, r.count - left, false});
                stack.push_back({r.x0, r.y0, r.x1, cut, left, false});
            } else {
                int cut = r.x0 + 1;
                for (int xx = r.x0; xx < r.x1 - 1; xx++) {
                    for (int yy = r.y0; yy < r.y1; yy++) acc += cost[(size_t)yy * gw + xx];
                    cut = xx + 1;
                    if (acc >= target) break;
                }
                stack.push_back({cut, r.y0, r.x1, r.y1, r.count - left, true});
                stack.push_back({r.x0, r.y0, cut, r.y1, left, true});
            }
        }
        out.resize((size_t)std::max(0, count), RenderTask{0, 0, 0, 0});  // workers beyond the number of blocks get empty tasks
        return out;
    }

    // LPT block order (RenderManager::lptFrame; sched.py lpt_block_order is the same function on the device): the 8x4 blocks of a
    // bw-wide grid, most expensive cost CLASS first — class = cost * levels / (max cost + 1) — and along the Z-order curve within a
    // class.  A full sort by cost scatters neighbouring blocks over the launch; classes keep the "long chains start first" property
    // of the sort and leave warps that run side by side on neighbouring pixels, i.e. on the same triangles (L1 hits): 728 -> 700 ms
    // per 1080p / 1024 spp frame on one GPU, 392 -> 374 ms on two (profiles/r02_block_order_ab.txt).
    static uint32_t lptLevels(int gpus) { return gpus <= 2 ? 8u : 4u; }
    static uint32_t zOrder(uint32_t x, uint32_t y) {
        auto spread = [](uint32_t v) {
            v &= 0xFFFFu;
            v = (v | (v << 8)) & 0x00FF00FFu;
            v = (v | (v << 4)) & 0x0F0F0F0Fu;
            v = (v | (v << 2)) & 0x33333333u;
            v = (v | (v << 1)) & 0x55555555u;
            return v;
        };
        return spread(x) | (spread(y) << 1);
    }
    static std::vector<uint32_t> lptBlockOrder(const uint32_t *costs, uint32_t n, uint32_t bw, uint32_t levels) {
        uint32_t cmax = 0;
        for (uint32_t i = 0; i < n; i++) cmax = std::max(cmax, costs[i]);
        std::vector<uint64_t> key(n);
        for (uint32_t i = 0; i < n; i++) {
            const uint64_t cls = (uint64_t)costs[i] * levels / ((uint64_t)cmax + 1);
            key[i] = ((levels - 1 - cls) << 32) | zOrder(i % bw, i / bw);  // ascending: highest class first, then along the curve
        }
        std::vector<uint32_t> order(n);
        for (uint32_t i = 0; i < n; i++) order[i] = i;
        std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return key[a] < key[b]; });  // keys are unique
        return order;
    }

    void setRes(int width, int height) {
        width_ = width;
        height_ = height;
    }

private:
    int width_ = 0;
    int height_ = 0;
};
