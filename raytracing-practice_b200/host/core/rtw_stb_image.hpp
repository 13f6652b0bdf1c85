// Include-layout shim: the reference's scene code includes "core/rtw_stb_image.hpp"
// (src/main.cpp:1-9).  Everything lives in rtb200_host.hpp.
#pragma once
#include "../rtb200_host.hpp"
