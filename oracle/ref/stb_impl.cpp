// stb_impl.cpp — instantiates the reference's vendored stb_image / stb_image_write exactly as the reference does in
// src/ui/display.cpp:3-7 (a UI translation unit that is otherwise out of scope). TEST INFRASTRUCTURE ONLY.
#define STB_IMAGE_WRITE_IMPLEMENTATION
#include <stb/stbi_image_write.h>

#define STB_IMAGE_IMPLEMENTATION
#include <stb/stb_image.h>
