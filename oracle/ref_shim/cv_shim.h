/* Minimal stand-in for the OpenCV 2.4.6 types the reference's LocalSearchKLT and PSN_* classes touch
 * (cv::Point, cv::Point2f and its operator-, cv::Point3d, cv::Rect, cv::vector).  TEST INFRASTRUCTURE ONLY:
 * it exists so that the reference's own source lines compile unchanged into oracle/_ref (see oracle/Makefile).
 * Semantics follow OpenCV's core/core.hpp: Point_<float> - Point_<float> is a float subtraction
 * (saturate_cast<float>(a.x - b.x)), Rect_<int> is {x, y, width, height}; in 2.4 `cv::vector` is std::vector. */
#pragma once
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <vector>

namespace cv {
using std::vector;
struct Point { int x, y; Point() : x(0), y(0) {} Point(int x_, int y_) : x(x_), y(y_) {} };
struct Point2f { float x, y; Point2f() : x(0), y(0) {} Point2f(float x_, float y_) : x(x_), y(y_) {} };
static inline Point2f operator-(const Point2f &a, const Point2f &b) { return Point2f((float)(a.x - b.x), (float)(a.y - b.y)); }
struct Point3d { double x, y, z; Point3d() : x(0), y(0), z(0) {} Point3d(double x_, double y_, double z_) : x(x_), y(y_), z(z_) {} };
struct Rect { int x, y, width, height; Rect() : x(0), y(0), width(0), height(0) {} Rect(int x_, int y_, int w_, int h_) : x(x_), y(y_), width(w_), height(h_) {} };
}  // namespace cv
