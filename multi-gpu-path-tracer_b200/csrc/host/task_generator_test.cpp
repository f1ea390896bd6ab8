// task_generator_test.cpp — the frame-to-task arithmetic of the host mirror without a GPU: equal tasks (reference
// src/Scheduling/TaskGenerator.h:46-80), DYNAMIC tiles, and the two time-driven schedulers DSFL / DSDL
// (reference src/RenderManager.h:264-408, :546-639).  Property under test everywhere: the tasks tile the frame exactly.
#include "TaskGenerator.h"

#include <cstdint>
#include <cstdio>
#include <random>
#include <vector>

static int failures = 0;
static void check(bool ok, const char *what) {
    if (!ok) {
        printf("FAIL %s\n", what);
        failures++;
    }
}

// every pixel covered exactly once, nothing outside the frame
static bool tilesExactly(const std::vector<RenderTask> &tasks, int W, int H) {
    std::vector<uint8_t> seen((size_t)W * H, 0);
    for (const RenderTask &t : tasks) {
        if (t.width < 0 || t.height < 0 || t.offset_x < 0 || t.offset_y < 0 || t.offset_x + t.width > W || t.offset_y + t.height > H) return false;
        for (int y = t.offset_y; y < t.offset_y + t.height; y++)
            for (int x = t.offset_x; x < t.offset_x + t.width; x++)
                if (seen[(size_t)y * W + x]++) return false;
    }
    for (uint8_t s : seen)
        if (s != 1) return false;
    return true;
}

static std::vector<std::vector<int>> layoutFor(int total, int maxInRow) {  // RenderManager::getTaskLayout
    std::vector<std::vector<int>> layout;
    int task = 0;
    while (task < total) {
        layout.push_back({});
        for (int r = 0; r < maxInRow && task < total; r++) layout.back().push_back(task++);
    }
    return layout;
}

int main() {
    TaskGenerator gen;
    std::mt19937 rng(1984);

    // equal columns and equal cells: exact cover, remainder in the last column / row (reference :46-55, :58-80)
    for (int n : {1, 2, 3, 7, 8}) {
        auto cols = gen.generateEqualTasks(n, 1921, 1080);
        check((int)cols.size() == n && tilesExactly(cols, 1921, 1080), "equal columns tile the frame");
        check(cols[0].width == 1921 / n && cols.back().width == 1921 - (n - 1) * (1921 / n), "the last column takes the remainder");
        for (int perRow : {1, 2, 3}) {
            auto layout = layoutFor(n, perRow);
            auto cells = gen.generateEqualTasks(n, layout, 1921, 1083);
            check((int)cells.size() == n && tilesExactly(cells, 1921, 1083), "equal cells tile the frame");
        }
    }
    // DYNAMIC tiles, ragged edges
    for (auto wh : {std::pair<int, int>{1920, 1080}, {37, 23}, {8, 4}, {1, 1}}) {
        auto tiles = gen.generateTiles(64, 32, wh.first, wh.second);
        check(tilesExactly(tiles, wh.first, wh.second), "tiles cover the frame");
        check((int)tiles.size() == ((wh.first + 63) / 64) * ((wh.second + 31) / 32), "tile count");
    }

    // DSFL: 400 frames of random times; the frame stays tiled, borders move by at most one thread block per frame
    for (int n : {2, 4, 6, 7}) {
        const int W = 640, H = 363, bx = 8, by = 8;
        auto layout = layoutFor(n, 2);
        auto tasks = gen.generateEqualTasks(n, layout, W, H);
        for (int frame = 0; frame < 400; frame++) {
            auto before = tasks;
            for (auto &t : tasks) t.time = 1 + (int)(rng() % 50);
            gen.adjustTasksDSFL(tasks, layout, W, H, bx, by);
            if (!tilesExactly(tasks, W, H)) { check(false, "DSFL keeps the frame tiled"); break; }
            bool small = true;
            for (size_t i = 0; i < tasks.size(); i++) {
                const bool lastInRow = tasks[i].offset_x + tasks[i].width == W, lastRow = tasks[i].offset_y + tasks[i].height == H;
                if (!lastInRow && std::abs(tasks[i].width - before[i].width) > bx) small = false;
                if (!lastRow && std::abs(tasks[i].height - before[i].height) > by) small = false;
            }
            check(small, "DSFL moves a border by at most one thread block per frame");
        }
    }
    // DSFL converges: left task four times as slow per pixel as the right one -> the border settles where the times are equal
    {
        const int W = 800, H = 64;
        auto layout = layoutFor(2, 2);
        auto tasks = gen.generateEqualTasks(2, layout, W, H);
        for (int frame = 0; frame < 300; frame++) {
            tasks[0].time = std::max(1, tasks[0].width * 4 / 10);
            tasks[1].time = std::max(1, tasks[1].width * 1 / 10);
            gen.adjustTasksDSFL(tasks, layout, W, H, 8, 8);
        }
        check(tilesExactly(tasks, W, H) && std::abs(tasks[0].width - W / 5) <= 16, "DSFL settles at the equal-time border (1/5 : 4/5)");
    }

    // DSDL: one rectangle per worker, exact cover, expensive regions get smaller rectangles
    for (int n : {1, 2, 4, 8}) {
        const int W = 640, H = 360;
        auto layout = layoutFor(n, 2);
        auto tasks = gen.generateEqualTasks(n, layout, W, H);
        for (int frame = 0; frame < 50; frame++) {
            for (auto &t : tasks) t.time = 1 + (int)(rng() % 90);
            tasks = gen.bisectTasksDSDL(tasks, n, W, H, 8, 8);
            check((int)tasks.size() == n && tilesExactly(tasks, W, H), "DSDL tiles the frame with one rectangle per worker");
        }
    }
    {
        const int W = 640, H = 360;
        auto tasks = gen.generateEqualTasks(2, 640, 360);  // two columns
        tasks[0].time = 90;                                  // the left half is nine times as expensive
        tasks[1].time = 10;
        auto out = gen.bisectTasksDSDL(tasks, 4, W, H, 8, 8);
        long leftArea = 0;
        int leftCount = 0;
        for (auto &t : out)
            if (t.offset_x + t.width <= W / 2) { leftArea += (long)t.width * t.height; leftCount++; }
        check(tilesExactly(out, W, H) && leftCount >= 2 && leftArea / std::max(1, leftCount) < (long)W * H / 4, "DSDL gives the expensive half more, smaller rectangles");
    }
    // more workers than thread blocks: the surplus gets empty tasks, the frame is still tiled
    {
        auto tasks = gen.generateEqualTasks(1, 8, 8);
        tasks[0].time = 5;
        auto out = gen.bisectTasksDSDL(tasks, 4, 8, 8, 8, 8);
        check((int)out.size() == 4 && tilesExactly(out, 8, 8), "DSDL with more workers than blocks");
    }
    // LPT block order: cost classes, highest first, Z-order within a class (the literal is the one tests/test_host_logic.py checks
    // sched.lpt_block_order against); random costs: a permutation with non-increasing classes and increasing curve index per class
    {
        const uint32_t costs[8] = {79, 10, 50, 30, 40, 20, 60, 0};
        const std::vector<uint32_t> want = {0, 6, 4, 2, 5, 3, 1, 7};
        check(TaskGenerator::lptBlockOrder(costs, 8, 4, 4) == want, "LPT block order of the 4x2 example");
        for (uint32_t levels : {1u, 4u, 8u}) {
            const uint32_t bw = 37, bh = 23, n = bw * bh;
            std::vector<uint32_t> c(n);
            uint32_t cmax = 0;
            for (auto &v : c) { v = rng() % 5000; cmax = std::max(cmax, v); }
            auto order = TaskGenerator::lptBlockOrder(c.data(), n, bw, levels);
            std::vector<int> seen(n, 0);
            bool ok = order.size() == n;
            for (uint32_t k = 0; ok && k < n; k++) {
                ok = order[k] < n && !seen[order[k]]++;
                if (ok && k) {
                    const uint64_t a = (uint64_t)c[order[k - 1]] * levels / ((uint64_t)cmax + 1), b = (uint64_t)c[order[k]] * levels / ((uint64_t)cmax + 1);
                    ok = a > b || (a == b && TaskGenerator::zOrder(order[k - 1] % bw, order[k - 1] / bw) < TaskGenerator::zOrder(order[k] % bw, order[k] / bw));
                }
            }
            check(ok, "LPT block order: permutation, classes descending, Z-order within a class");
        }
        check(TaskGenerator::lptBlockOrder(nullptr, 0, 1, 4).empty(), "LPT block order of an empty grid");
    }
    printf("%s\n", failures ? "TASK_GENERATOR_TEST_FAILED" : "TASK_GENERATOR_TEST_OK");
    return failures ? 1 : 0;
}
