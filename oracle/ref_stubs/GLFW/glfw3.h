/* oracle/ref_stubs: OR/Camera.h:42 names GLFWwindow (the viewer is not compiled) */
#pragma once
typedef struct GLFWwindow GLFWwindow;
