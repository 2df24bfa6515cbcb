/* oracle/ref_stubs: the GL types OR/shaderClass.h and OR/Camera.h name (the viewer is not compiled) */
#pragma once
typedef unsigned int GLuint; typedef int GLint; typedef unsigned int GLenum; typedef float GLfloat; typedef int GLsizei; typedef char GLchar;
