// cvshim.hpp -- TEST INFRASTRUCTURE (oracle/_ref build only).
//
// A minimal, self-written stand-in for the OpenCV C++ headers that the reference's vendored line_descriptor sources
// (/root/reference/src/line_descriptor/src/*.cpp) include, so that those sources compile UNMODIFIED, from where they
// lie, without OpenCV's C++ development files (which this image does not have).  Only what the descriptor path
// executes is functional:
//   cv::Mat (ref-counted 2-D buffer), Point/Size/Vec, Ptr, LineIterator::count, cvtColor(BGR2GRAY),
//   GaussianBlur(5x5, sigma 1), Sobel(3x3 -> CV_16S)   -- these three call the C oracle (oracle/csrc/lane_oracle.c),
//   whose arithmetic is pinned bit-exactly to cv2 4.13 by tests/test_oracle.py;
//   cv::LineSegmentDetector::detect returns the segments the harness injected (the reference's LSDDetectorC then
//   fills its KeyLines from them).
// Everything else (EDLines, drawing, FileStorage, pyrDown, resize ...) is declared so the sources compile and throws
// std::logic_error if ever called.
#pragma once
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#define CV_EXPORTS
#define CV_EXPORTS_W
#define CV_WRAP
#define CV_OUT
#define CV_IN_OUT
#define CV_PROP
#define CV_PROP_RW

typedef unsigned char uchar;
typedef signed char schar;
typedef unsigned short ushort;
typedef int64_t int64;
typedef uint64_t uint64;

#define CV_CN_SHIFT 3
#define CV_8U 0
#define CV_8S 1
#define CV_16U 2
#define CV_16S 3
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn) - 1) << CV_CN_SHIFT))
#define CV_MAT_DEPTH(t) ((t) & 7)
#define CV_MAT_CN(t) ((((t) >> CV_CN_SHIFT) & 63) + 1)
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_8SC1 CV_MAKETYPE(CV_8S, 1)
#define CV_16UC1 CV_MAKETYPE(CV_16U, 1)
#define CV_16SC1 CV_MAKETYPE(CV_16S, 1)
#define CV_32SC1 CV_MAKETYPE(CV_32S, 1)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)
#define CV_PI 3.1415926535897932384626433832795
#define CV_Assert(e) do { if (!(e)) throw std::runtime_error("CV_Assert failed: " #e); } while (0)
#define CV_Error(code, msg) throw std::runtime_error(std::string(msg))
#define CV_StsBadArg -5
#define CV_StsNotImplemented -213

extern "C" {
// oracle/csrc/lane_oracle.c (pinned to cv2 4.13)
void orc_gauss5_sobel(const uint8_t *gray, int H, int W, uint8_t *blur, int16_t *dx, int16_t *dy);
void orc_bgr2gray(const uint8_t *bgr, size_t n, uint8_t *gray);
}

inline int cvRound(double v) { return (int)lrint(v); }      // round half to even, like OpenCV's SSE2 path
inline int cvFloor(double v) { int i = (int)v; return i - (i > v); }
inline int cvCeil(double v) { int i = (int)v; return i + (i < v); }

namespace cv {

using std::abs;
using std::exp;
using std::log;
using std::max;
using std::min;
using std::pow;
using std::sqrt;
using std::swap;

typedef std::string String;

[[noreturn]] inline void shim_unimplemented(const char *what)
{
    throw std::logic_error(std::string("cvshim: ") + what + " is not implemented (not on the descriptor path)");
}

template <typename T> inline T saturate_cast(double v) { return (T)v; }
template <> inline int saturate_cast<int>(double v) { return cvRound(v); }
template <> inline uchar saturate_cast<uchar>(double v) { int i = cvRound(v); return (uchar)(i < 0 ? 0 : i > 255 ? 255 : i); }
template <> inline short saturate_cast<short>(double v) { int i = cvRound(v); return (short)(i < -32768 ? -32768 : i > 32767 ? 32767 : i); }

// ---- small value types ---------------------------------------------------------------------------------------
template <typename T> struct Size_ {
    T width, height;
    Size_() : width(0), height(0) {}
    Size_(T w, T h) : width(w), height(h) {}
    bool operator==(const Size_ &o) const { return width == o.width && height == o.height; }
    bool operator!=(const Size_ &o) const { return !(*this == o); }
    T area() const { return width * height; }
};
typedef Size_<int> Size;

template <typename T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T x_, T y_) : x(x_), y(y_) {}
    template <typename U> Point_(const Point_<U> &p) : x(saturate_cast<T>(p.x)), y(saturate_cast<T>(p.y)) {}
    Point_ operator+(const Point_ &o) const { return Point_(x + o.x, y + o.y); }
    Point_ operator-(const Point_ &o) const { return Point_(x - o.x, y - o.y); }
    bool operator==(const Point_ &o) const { return x == o.x && y == o.y; }
};
template <> template <> inline Point_<float>::Point_(const Point_<int> &p) : x((float)p.x), y((float)p.y) {}
typedef Point_<int> Point;
typedef Point_<float> Point2f;
typedef Point_<double> Point2d;

template <typename T> struct Rect_ {
    T x, y, width, height;
    Rect_() : x(0), y(0), width(0), height(0) {}
    Rect_(T a, T b, T c, T d) : x(a), y(b), width(c), height(d) {}
};
typedef Rect_<int> Rect;

template <typename T, int N> struct Vec {
    T val[N];
    Vec() { for (int i = 0; i < N; ++i) val[i] = T(); }
    Vec(T a, T b) { static_assert(N >= 2, ""); val[0] = a; val[1] = b; for (int i = 2; i < N; ++i) val[i] = T(); }
    Vec(T a, T b, T c) { static_assert(N >= 3, ""); val[0] = a; val[1] = b; val[2] = c; for (int i = 3; i < N; ++i) val[i] = T(); }
    Vec(T a, T b, T c, T d) { static_assert(N >= 4, ""); val[0] = a; val[1] = b; val[2] = c; val[3] = d; for (int i = 4; i < N; ++i) val[i] = T(); }
    T &operator[](int i) { return val[i]; }
    const T &operator[](int i) const { return val[i]; }
};
typedef Vec<float, 4> Vec4f;
typedef Vec<int, 4> Vec4i;
typedef Vec<int, 3> Vec3i;
typedef Vec<uchar, 3> Vec3b;
typedef Vec<float, 2> Vec2f;

struct Scalar {
    double val[4];
    Scalar(double a = 0, double b = 0, double c = 0, double d = 0) { val[0] = a; val[1] = b; val[2] = c; val[3] = d; }
    static Scalar all(double v) { return Scalar(v, v, v, v); }
    double &operator[](int i) { return val[i]; }
    const double &operator[](int i) const { return val[i]; }
    bool operator==(const Scalar &o) const { return !memcmp(val, o.val, sizeof(val)); }
};

struct Range { int start, end; Range() : start(0), end(0) {} Range(int s, int e) : start(s), end(e) {} };

template <typename T> class Ptr : public std::shared_ptr<T> {
public:
    Ptr() {}
    Ptr(T *p) : std::shared_ptr<T>(p) {}
    template <typename U> Ptr(const std::shared_ptr<U> &o) : std::shared_ptr<T>(o) {}
    bool empty() const { return !this->get(); }
    operator T *() const { return this->get(); }
};
template <typename T, typename... A> Ptr<T> makePtr(A &&...a) { return Ptr<T>(new T(std::forward<A>(a)...)); }

// ---- Mat ---------------------------------------------------------------------------------------------------------
inline size_t shim_elem_size(int type)
{
    static const int sz[8] = {1, 1, 2, 2, 4, 4, 8, 2};
    return (size_t)sz[CV_MAT_DEPTH(type)] * CV_MAT_CN(type);
}

class MatExpr;

class Mat {
public:
    int flags;          // = type
    int rows, cols;
    uchar *data;
    size_t step;
    std::shared_ptr<std::vector<uchar> > buf;

    Mat() : flags(0), rows(0), cols(0), data(nullptr), step(0) {}
    Mat(int r, int c, int type) : Mat() { create(r, c, type); }
    Mat(Size s, int type) : Mat() { create(s.height, s.width, type); }
    Mat(int r, int c, int type, const Scalar &s) : Mat() { create(r, c, type); setTo(s); }
    Mat(Size sz, int type, const Scalar &s) : Mat() { create(sz.height, sz.width, type); setTo(s); }
    Mat(int r, int c, int type, void *ext, size_t st = 0) : flags(type), rows(r), cols(c), data((uchar *)ext), step(st ? st : c * shim_elem_size(type)) {}
    Mat(const MatExpr &e);
    Mat &operator=(const MatExpr &e);

    void create(int r, int c, int type)
    {
        if (data && rows == r && cols == c && flags == type && buf) return;
        flags = type; rows = r; cols = c; step = (size_t)c * shim_elem_size(type);
        buf = std::make_shared<std::vector<uchar> >((size_t)r * step + 16);
        data = buf->data();
    }
    void create(Size s, int type) { create(s.height, s.width, type); }
    void release() { buf.reset(); data = nullptr; rows = cols = 0; step = 0; }
    Mat clone() const
    {
        Mat m;
        if (!data) return m;
        m.create(rows, cols, flags);
        for (int y = 0; y < rows; ++y) memcpy(m.data + (size_t)y * m.step, data + (size_t)y * step, (size_t)cols * elemSize());
        return m;
    }
    void copyTo(Mat &dst) const { dst = clone(); }
    void copyTo(const class _OutputArray &dst) const;
    class MatExpr t() const;
    class MatExpr inv(int = 0) const;
    void convertTo(Mat &, int, double = 1, double = 0) const { shim_unimplemented("Mat::convertTo"); }
    int type() const { return flags; }
    int depth() const { return CV_MAT_DEPTH(flags); }
    int channels() const { return CV_MAT_CN(flags); }
    size_t elemSize() const { return shim_elem_size(flags); }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    size_t total() const { return (size_t)rows * cols; }
    Size size() const { return Size(cols, rows); }
    bool isContinuous() const { return step == (size_t)cols * elemSize(); }
    template <typename T = uchar> T *ptr(int y = 0) { return (T *)(data + (size_t)y * step); }
    template <typename T = uchar> const T *ptr(int y = 0) const { return (const T *)(data + (size_t)y * step); }
    template <typename T> T &at(int y, int x) { return ((T *)(data + (size_t)y * step))[x]; }
    template <typename T> const T &at(int y, int x) const { return ((const T *)(data + (size_t)y * step))[x]; }
    template <typename T> T &at(Point p) { return at<T>(p.y, p.x); }
    template <typename T> const T &at(Point p) const { return at<T>(p.y, p.x); }
    template <typename T> T &at(int i) { return rows == 1 ? at<T>(0, i) : at<T>(i, 0); }
    template <typename T> const T &at(int i) const { return rows == 1 ? at<T>(0, i) : at<T>(i, 0); }
    Mat row(int y) const { Mat m = *this; m.rows = 1; m.data = data + (size_t)y * step; return m; }
    Mat col(int) const { shim_unimplemented("Mat::col"); }
    Mat &setTo(const Scalar &s)
    {
        for (int y = 0; y < rows; ++y)
            for (int x = 0; x < cols * channels(); ++x) {
                double v = s.val[x % channels()];
                switch (depth()) {
                case CV_8U: ptr<uchar>(y)[x] = saturate_cast<uchar>(v); break;
                case CV_8S: ptr<schar>(y)[x] = (schar)v; break;
                case CV_16S: ptr<short>(y)[x] = saturate_cast<short>(v); break;
                case CV_16U: ptr<ushort>(y)[x] = (ushort)v; break;
                case CV_32S: ptr<int>(y)[x] = (int)v; break;
                case CV_32F: ptr<float>(y)[x] = (float)v; break;
                default: ptr<double>(y)[x] = v; break;
                }
            }
        return *this;
    }
    Mat &operator=(const Scalar &s) { return setTo(s); }
    static Mat zeros(int r, int c, int type) { return Mat(r, c, type, Scalar(0)); }
    static Mat zeros(Size s, int type) { return Mat(s, type, Scalar(0)); }
    static Mat ones(int r, int c, int type) { return Mat(r, c, type, Scalar(1)); }
    static Mat ones(Size s, int type) { return Mat(s, type, Scalar(1)); }
    void push_back(const Mat &) { shim_unimplemented("Mat::push_back"); }
};

// expressions never run on the descriptor path: enough surface to compile
class MatExpr {
public:
    Mat m;
    MatExpr() {}
    MatExpr(const Mat &a) : m(a) {}
};
inline Mat::Mat(const MatExpr &e) : Mat(e.m) {}
inline MatExpr Mat::t() const { shim_unimplemented("Mat::t"); }
inline MatExpr Mat::inv(int) const { shim_unimplemented("Mat::inv"); }
inline MatExpr operator*(const MatExpr &, const MatExpr &) { shim_unimplemented("MatExpr * MatExpr"); }
inline MatExpr operator*(const Mat &, const MatExpr &) { shim_unimplemented("Mat * MatExpr"); }
inline MatExpr operator*(const MatExpr &, const Mat &) { shim_unimplemented("MatExpr * Mat"); }
inline Mat &Mat::operator=(const MatExpr &e) { *this = e.m; return *this; }
inline MatExpr operator*(const Mat &, double) { shim_unimplemented("Mat * scalar"); }
inline MatExpr operator*(double, const Mat &) { shim_unimplemented("scalar * Mat"); }
inline MatExpr operator*(const Mat &, const Mat &) { shim_unimplemented("Mat * Mat"); }
inline MatExpr operator+(const Mat &, const Mat &) { shim_unimplemented("Mat + Mat"); }
inline MatExpr operator-(const Mat &, const Mat &) { shim_unimplemented("Mat - Mat"); }
inline MatExpr operator/(const Mat &, double) { shim_unimplemented("Mat / scalar"); }
inline MatExpr abs(const Mat &) { shim_unimplemented("abs(Mat)"); }

template <typename T> struct shim_type;
template <> struct shim_type<uchar> { enum { value = CV_8UC1 }; };
template <> struct shim_type<schar> { enum { value = CV_8SC1 }; };
template <> struct shim_type<short> { enum { value = CV_16SC1 }; };
template <> struct shim_type<ushort> { enum { value = CV_16UC1 }; };
template <> struct shim_type<int> { enum { value = CV_32SC1 }; };
template <> struct shim_type<float> { enum { value = CV_32FC1 }; };
template <> struct shim_type<double> { enum { value = CV_64FC1 }; };

template <typename T> class Mat_ : public Mat {
public:
    Mat_() { flags = shim_type<T>::value; }
    Mat_(int r, int c) : Mat(r, c, shim_type<T>::value) {}
    Mat_(int r, int c, const T &v) : Mat(r, c, shim_type<T>::value, Scalar((double)v)) {}
    Mat_(const Mat &m) : Mat(m) {}
    Mat_(const MatExpr &e) : Mat(e) {}
    T &operator()(int y, int x) { return this->template at<T>(y, x); }
    const T &operator()(int y, int x) const { return this->template at<T>(y, x); }
    T *operator[](int y) { return this->template ptr<T>(y); }
    const T *operator[](int y) const { return this->template ptr<T>(y); }
    static Mat_ zeros(int r, int c) { return Mat_(r, c, T(0)); }
    static Mat_ ones(int r, int c) { return Mat_(r, c, T(1)); }
};

// ---- array arguments ---------------------------------------------------------------------------------------------------
class _InputArray {
public:
    const Mat *m;
    const std::vector<Mat> *vm;
    Mat own;
    _InputArray() : m(nullptr), vm(nullptr) {}
    _InputArray(const Mat &a) : m(&a), vm(nullptr) {}
    _InputArray(const MatExpr &e) : m(nullptr), vm(nullptr), own(e.m) { m = &own; }
    _InputArray(const std::vector<Mat> &v) : m(nullptr), vm(&v) {}
    template <typename T> _InputArray(const std::vector<T> &v) : m(nullptr), vm(nullptr), own(1, (int)v.size(), CV_8UC1, (void *)v.data()) { m = &own; }
    Mat getMat(int i = -1) const { return i >= 0 && vm ? (*vm)[i] : (m ? *m : Mat()); }
    void getMatVector(std::vector<Mat> &out) const { if (vm) out = *vm; else out.assign(1, getMat()); }
    bool empty() const { return vm ? vm->empty() : (!m || m->empty()); }
    Size size() const { return getMat().size(); }
    int type() const { return getMat().type(); }
};
class _OutputArray : public _InputArray {
public:
    Mat *om;
    std::vector<Mat> *ovm;
    _OutputArray() : om(nullptr), ovm(nullptr) {}
    _OutputArray(Mat &a) : _InputArray(a), om(&a), ovm(nullptr) {}
    _OutputArray(std::vector<Mat> &v) : _InputArray(v), om(nullptr), ovm(&v) {}
    void create(int r, int c, int type) const { if (om) om->create(r, c, type); }
    void create(Size s, int type) const { if (om) om->create(s, type); }
    Mat &getMatRef() const { return *om; }
    bool needed() const { return om || ovm; }
};
inline void Mat::copyTo(const _OutputArray &dst) const { if (dst.om) *dst.om = clone(); }
typedef const _InputArray &InputArray;
typedef const _InputArray &InputArrayOfArrays;
typedef const _OutputArray &OutputArray;
typedef const _OutputArray &OutputArrayOfArrays;
typedef const _OutputArray &InputOutputArray;
inline const _OutputArray &noArray() { static _OutputArray none; return none; }

// ---- persistence / algorithm stubs -----------------------------------------------------------------------------------
class FileNode {
public:
    FileNode operator[](const char *) const { return FileNode(); }
    FileNode operator[](const String &) const { return FileNode(); }
    bool empty() const { return true; }
    operator int() const { return 0; }
    operator float() const { return 0; }
    operator double() const { return 0; }
};
template <typename T> inline void operator>>(const FileNode &, T &) {}
class FileStorage {
public:
    enum { READ = 0, WRITE = 1 };
    FileStorage() {}
    FileStorage(const String &, int) {}
    bool isOpened() const { return false; }
    FileNode root() const { return FileNode(); }
    FileNode getFirstTopLevelNode() const { return FileNode(); }
    FileNode operator[](const char *) const { return FileNode(); }
    void release() {}
};
template <typename T> inline FileStorage &operator<<(FileStorage &fs, const T &) { return fs; }

class Algorithm {
public:
    virtual ~Algorithm() {}
    virtual void clear() {}
    virtual void read(const FileNode &) {}
    virtual void write(FileStorage &) const {}
    virtual bool empty() const { return false; }
    virtual void save(const String &) const {}
};

struct KeyPoint {
    Point2f pt; float size, angle, response; int octave, class_id;
    KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) {}
};
struct DMatch {
    int queryIdx, trainIdx, imgIdx; float distance;
    DMatch() : queryIdx(-1), trainIdx(-1), imgIdx(-1), distance(3.402823466e+38f) {}
    DMatch(int q, int t, float d) : queryIdx(q), trainIdx(t), imgIdx(-1), distance(d) {}
    DMatch(int q, int t, int i, float d) : queryIdx(q), trainIdx(t), imgIdx(i), distance(d) {}
    bool operator<(const DMatch &m) const { return distance < m.distance; }
};

class RNG {
public:
    uint64 state;
    RNG(uint64 s = 0xffffffff) : state(s ? s : 0xffffffff) {}
    unsigned next() { state = (uint64)(unsigned)state * 4164903690U + (unsigned)(state >> 32); return (unsigned)state; }
    int uniform(int a, int b) { return a == b ? a : (int)(next() % (unsigned)(b - a) + a); }
    operator unsigned() { return next(); }
    unsigned operator()(unsigned n) { return next() % n; }
};
inline RNG &theRNG() { static RNG r; return r; }

// ---- imgproc ------------------------------------------------------------------------------------------------------
enum { COLOR_BGR2GRAY = 6, COLOR_GRAY2BGR = 8, COLOR_RGB2GRAY = 7 };
enum { INTER_NEAREST = 0, INTER_LINEAR = 1, INTER_CUBIC = 2, INTER_AREA = 3, INTER_LINEAR_EXACT = 5 };
enum { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1, BORDER_REFLECT = 2, BORDER_REFLECT_101 = 4, BORDER_DEFAULT = 4 };
enum { THRESH_BINARY = 0, THRESH_BINARY_INV = 1, THRESH_TRUNC = 2, THRESH_TOZERO = 3, THRESH_TOZERO_INV = 4 };
enum { CMP_EQ = 0, CMP_GT = 1, CMP_GE = 2, CMP_LT = 3, CMP_LE = 4, CMP_NE = 5 };
enum { LSD_REFINE_NONE = 0, LSD_REFINE_STD = 1, LSD_REFINE_ADV = 2 };
enum { LINE_8 = 8, LINE_AA = 16, FILLED = -1 };
enum { NORM_L2 = 4, NORM_HAMMING = 6 };

inline void cvtColor(InputArray src_, OutputArray dst_, int code, int = 0)
{
    Mat src = src_.getMat();
    if (code != COLOR_BGR2GRAY || src.type() != CV_8UC3) shim_unimplemented("cvtColor other than 8-bit BGR2GRAY");
    Mat s = src.isContinuous() ? src : src.clone();
    Mat out(src.rows, src.cols, CV_8UC1);
    orc_bgr2gray(s.data, s.total(), out.data);
    dst_.getMatRef() = out;
}

// cv::GaussianBlur(src, dst, Size(5,5), 1): the only call on the path (binary_descriptor_custom.cpp:358)
inline void GaussianBlur(InputArray src_, OutputArray dst_, Size k, double sx, double sy = 0, int border = BORDER_DEFAULT)
{
    Mat src = src_.getMat();
    if (k.width != 5 || k.height != 5 || sx != 1 || sy != 0 || border != BORDER_DEFAULT || src.type() != CV_8UC1)
        shim_unimplemented("GaussianBlur other than 8-bit 5x5 sigma 1");
    Mat s = src.isContinuous() ? src : src.clone();
    Mat out(src.rows, src.cols, CV_8UC1);
    std::vector<int16_t> dx(src.total()), dy(src.total());
    orc_gauss5_sobel(s.data, s.rows, s.cols, out.data, dx.data(), dy.data());
    dst_.getMatRef() = out;
}

// cv::Sobel(src, dst, CV_16SC1, dx, dy, 3) of an 8-bit image (binary_descriptor_custom.cpp:395-396).  The C oracle
// computes blur+Sobel in one call; its Sobel stage is applied here to an image that is ALREADY blurred, so the blur
// is undone by feeding the image through a pure 3x3 Sobel written out below (reflect-101 border, like cv2).
inline void Sobel(InputArray src_, OutputArray dst_, int ddepth, int dx, int dy, int ksize = 3, double scale = 1, double delta = 0,
                  int border = BORDER_DEFAULT)
{
    Mat src = src_.getMat();
    if (ddepth != CV_16SC1 || ksize != 3 || scale != 1 || delta != 0 || border != BORDER_DEFAULT || src.type() != CV_8UC1 ||
        dx + dy != 1)
        shim_unimplemented("Sobel other than 8-bit -> 16S, 3x3, first order");
    const int H = src.rows, W = src.cols;
    Mat out(H, W, CV_16SC1);
    auto R = [](int i, int n) { if (n == 1) return 0; if (i < 0) i = -i; if (i >= n) i = 2 * n - 2 - i; return i; };
    for (int y = 0; y < H; ++y) {
        const uchar *r0 = src.ptr<uchar>(R(y - 1, H)), *r1 = src.ptr<uchar>(y), *r2 = src.ptr<uchar>(R(y + 1, H));
        short *o = out.ptr<short>(y);
        for (int x = 0; x < W; ++x) {
            const int xm = R(x - 1, W), xp = R(x + 1, W);
            o[x] = dx ? (short)((r0[xp] + 2 * r1[xp] + r2[xp]) - (r0[xm] + 2 * r1[xm] + r2[xm]))
                      : (short)((r2[xm] + 2 * r2[x] + r2[xp]) - (r0[xm] + 2 * r0[x] + r0[xp]));
        }
    }
    dst_.getMatRef() = out;
}

inline void resize(InputArray, OutputArray, Size, double = 0, double = 0, int = INTER_LINEAR) { shim_unimplemented("resize"); }
inline void pyrDown(InputArray, OutputArray, const Size & = Size(), int = BORDER_DEFAULT) { shim_unimplemented("pyrDown"); }
inline double threshold(InputArray, OutputArray, double, double, int) { shim_unimplemented("threshold"); }
inline void compare(InputArray, InputArray, OutputArray, int) { shim_unimplemented("compare"); }
inline void add(InputArray, InputArray, OutputArray, InputArray = noArray(), int = -1) { shim_unimplemented("add"); }
inline void multiply(InputArray, InputArray, OutputArray, double = 1, int = -1) { shim_unimplemented("multiply"); }
inline void magnitude(InputArray, InputArray, OutputArray) { shim_unimplemented("magnitude"); }
inline void phase(InputArray, InputArray, OutputArray, bool = false) { shim_unimplemented("phase"); }
inline void flip(InputArray, OutputArray, int) { shim_unimplemented("flip"); }
inline Scalar mean(InputArray, InputArray = noArray()) { shim_unimplemented("mean"); }
inline Scalar sum(InputArray) { shim_unimplemented("sum"); }
inline double norm(InputArray, int = NORM_L2, InputArray = noArray()) { shim_unimplemented("norm"); }
inline double norm(InputArray, InputArray, int = NORM_L2, InputArray = noArray()) { shim_unimplemented("norm"); }
inline int countNonZero(InputArray) { shim_unimplemented("countNonZero"); }
inline float fastAtan2(float y, float x) { float a = (float)(atan2((double)y, (double)x) * 180.0 / CV_PI); return a < 0 ? a + 360.f : a; }
inline void line(Mat &, Point, Point, const Scalar &, int = 1, int = 8, int = 0) { shim_unimplemented("line"); }
inline void circle(Mat &, Point, int, const Scalar &, int = 1, int = 8, int = 0) { shim_unimplemented("circle"); }
inline void imshow(const String &, InputArray) { shim_unimplemented("imshow"); }
inline int waitKey(int = 0) { shim_unimplemented("waitKey"); }
inline int64 getTickCount() { return 0; }
inline double getTickFrequency() { return 1; }

// LineIterator: only `count` is read (LSDDetector_custom.cpp:187-188).  8-connected: max(|dx|, |dy|) + 1 after clipping
// the segment to the image; the endpoints arrive as Point (Point2f -> Point rounds with cvRound).
class LineIterator {
public:
    int count;
    LineIterator(const Mat &img, Point p1, Point p2, int connectivity = 8, bool = false)
    {
        const int W = img.cols, H = img.rows;
        auto inside = [&](Point p) { return p.x >= 0 && p.x < W && p.y >= 0 && p.y < H; };
        if (!inside(p1) || !inside(p2)) shim_unimplemented("LineIterator with endpoints outside the image (clipLine)");
        const int dx = std::abs(p2.x - p1.x), dy = std::abs(p2.y - p1.y);
        count = connectivity == 8 ? std::max(dx, dy) + 1 : dx + dy + 1;
    }
};

// ---- LineSegmentDetector: returns what the harness injected ------------------------------------------------------------
class LineSegmentDetector : public Algorithm {
public:
    static std::vector<Vec4f> &injected() { static std::vector<Vec4f> v; return v; }
    virtual void detect(InputArray, std::vector<Vec4f> &lines) { lines = injected(); }
    virtual void detect(InputArray, OutputArray, OutputArray = noArray(), OutputArray = noArray(), OutputArray = noArray())
    {
        shim_unimplemented("LineSegmentDetector::detect(OutputArray)");
    }
};
inline Ptr<LineSegmentDetector> createLineSegmentDetector(int = LSD_REFINE_STD, double = 0.8, double = 0.6, double = 2.0, double = 22.5,
                                                          double = 0, double = 0.7, int = 1024)
{
    return makePtr<LineSegmentDetector>();
}

}  // namespace cv
