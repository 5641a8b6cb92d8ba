/*
 * Headless stand-in for the reference's CameraController (/root/reference/src/utils/camera_controller.cpp:30-84):
 * only the start-up pose, which is the benchmark camera of every BASELINE config.  The
 * mouse/keyboard fly controls need a window and are out of scope.  sin/cos come from include/rt_math.h
 * (correctly rounded), so the struct is the same on every platform.
 */
#pragma once

#include <cmath>

#include "rt_math.h"
#include "types.hpp"

namespace rt_host
{

class CameraController
{
public:
    CameraController(std::uint32_t width, std::uint32_t height)
    {
        const float MATH_PIDIV2 = 1.570796327f;
        pitch_ = MATH_PIDIV2; yaw_ = MATH_PIDIV2;
        camera_data_ = {};
        camera_data_.focus_distance = 10.0f;
        camera_data_.position = make_float3(0.0f, -1.0f, 1.0f);
        camera_data_.fov = 75.0f * 3.1415f / 180.0f;
        camera_data_.aspect_ratio = (float)width / (float)height;
        Update();
    }

    // camera_controller.cpp:77-80 with dt = 0 (no input)
    void Update()
    {
        camera_data_.front = make_float3(rt_cosf(yaw_) * rt_sinf(pitch_), rt_sinf(yaw_) * rt_sinf(pitch_), rt_cosf(pitch_));
        float3 up = make_float3(0.0f, 0.0f, 1.0f);
        float3 r = cross(camera_data_.front, up);
        float len = std::sqrt(r.x * r.x + r.y * r.y + r.z * r.z);
        float3 right = make_float3(r.x / len, r.y / len, r.z / len);
        camera_data_.up = cross(right, camera_data_.front);
    }

    Camera const& GetData() const { return camera_data_; }
    void SetAperture(float a) { camera_data_.aperture = a; }
    void SetFocusDistance(float d) { camera_data_.focus_distance = d; }
    void SetPosition(float3 p) { camera_data_.position = p; }
    void SetData(Camera const& c) { camera_data_ = c; }     // headless drivers place the camera directly (no fly controls)

private:
    Camera camera_data_;
    float pitch_, yaw_;
};

} // namespace rt_host
