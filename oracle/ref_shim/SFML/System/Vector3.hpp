/* TEST INFRASTRUCTURE: stand-in for <SFML/System/Vector3.hpp> (SFML is not installed in this image), just enough for
 * the reference's include/util.hpp and src/map/Octree.cpp to compile from where they lie.  Own code, not SFML's. */
#pragma once
namespace sf {
template <typename T>
struct Vector3 {
    T x, y, z;
    Vector3() : x(0), y(0), z(0) {}
    Vector3(T X, T Y, T Z) : x(X), y(Y), z(Z) {}
    template <typename U>
    explicit Vector3(const Vector3<U> &v) : x(static_cast<T>(v.x)), y(static_cast<T>(v.y)), z(static_cast<T>(v.z)) {}
};
template <typename T> Vector3<T> operator-(const Vector3<T> &a) { return Vector3<T>(-a.x, -a.y, -a.z); }
template <typename T> Vector3<T> operator+(const Vector3<T> &a, const Vector3<T> &b) { return Vector3<T>(a.x + b.x, a.y + b.y, a.z + b.z); }
template <typename T> Vector3<T> operator-(const Vector3<T> &a, const Vector3<T> &b) { return Vector3<T>(a.x - b.x, a.y - b.y, a.z - b.z); }
template <typename T> Vector3<T> operator*(const Vector3<T> &a, T s) { return Vector3<T>(a.x * s, a.y * s, a.z * s); }
template <typename T> Vector3<T> operator*(T s, const Vector3<T> &a) { return Vector3<T>(a.x * s, a.y * s, a.z * s); }
template <typename T> Vector3<T> operator/(const Vector3<T> &a, T s) { return Vector3<T>(a.x / s, a.y / s, a.z / s); }
template <typename T> Vector3<T> &operator+=(Vector3<T> &a, const Vector3<T> &b) { a.x += b.x; a.y += b.y; a.z += b.z; return a; }
template <typename T> Vector3<T> &operator-=(Vector3<T> &a, const Vector3<T> &b) { a.x -= b.x; a.y -= b.y; a.z -= b.z; return a; }
template <typename T> Vector3<T> &operator*=(Vector3<T> &a, T s) { a.x *= s; a.y *= s; a.z *= s; return a; }
template <typename T> Vector3<T> &operator/=(Vector3<T> &a, T s) { a.x /= s; a.y /= s; a.z /= s; return a; }
template <typename T> bool operator==(const Vector3<T> &a, const Vector3<T> &b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
template <typename T> bool operator!=(const Vector3<T> &a, const Vector3<T> &b) { return !(a == b); }
typedef Vector3<int> Vector3i;
typedef Vector3<float> Vector3f;
}
