// Empty stand-in: the reference includes glm in HostScene.h but the hot path never uses it.
#pragma once
