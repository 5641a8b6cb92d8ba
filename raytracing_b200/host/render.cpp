#include "render.hpp"

#include <stdexcept>

namespace rt_host
{

Render::Render(std::uint32_t width, std::uint32_t height, RenderBackend backend, Scene& scene, const char* env_map_path, int device)
    : scene_(scene), width_(width), height_(height)
{
    Init(backend, env_map_path, nullptr, nullptr, device);
}

Render::Render(std::uint32_t width, std::uint32_t height, RenderBackend backend, Scene& scene, const char* env_map_path, std::vector<int> const& devices)
    : scene_(scene), width_(width), height_(height)
{
    Init(backend, env_map_path, nullptr, &devices, 0);
}

Render::Render(std::uint32_t width, std::uint32_t height, RenderBackend backend, Scene& scene, Image const& env_image, std::vector<int> const& devices)
    : scene_(scene), width_(width), height_(height)
{
    Init(backend, nullptr, &env_image, &devices, 0);
}

void Render::Init(RenderBackend backend, const char* env_map_path, const Image* env_image, const std::vector<int>* devices, int device)
{
    if (backend != RenderBackend::kCUDA)
        throw std::runtime_error("this build only provides RenderBackend::kCUDA (the OpenCL/OpenGL backends live in the reference)");
    camera_controller_ = std::make_unique<CameraController>(width_, height_);
    // render.cpp:60-67: build the BVH (reorders the scene's triangles), THEN finalize the scene
    acc_structure_ = std::make_unique<Bvh>();
    acc_structure_->BuildCPU(scene_.GetTriangles());
    if (env_image) scene_.Finalize(env_image->data.data(), env_image->width, env_image->height);
    else scene_.Finalize(env_map_path);
    std::unique_ptr<CUDAPathTraceIntegrator> cuda;
    if (devices) cuda = std::make_unique<CUDAPathTraceIntegrator>(width_, height_, *acc_structure_, *devices);
    else cuda = std::make_unique<CUDAPathTraceIntegrator>(width_, height_, *acc_structure_, device);
    image_.assign((size_t)width_ * height_ * 4, 0.0f);
    cuda->SetResolveTarget(image_.data(), true);             // page-locked: device->host copies are asynchronous
    integrator_ = std::move(cuda);
    integrator_->UploadGPUData(scene_, *acc_structure_);     // render.cpp:82
}

// render.cpp:172-204 without window, GUI and presentation
void Render::RenderFrame()
{
    integrator_->SetCameraData(camera_controller_->GetData());
    if (camera_changed_)
    {
        integrator_->RequestReset();
        camera_changed_ = false;
    }
    integrator_->Integrate();
}

} // namespace rt_host
