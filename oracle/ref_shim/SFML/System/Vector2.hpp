/* TEST INFRASTRUCTURE: stand-in for <SFML/System/Vector2.hpp>; see Vector3.hpp. */
#pragma once
namespace sf {
template <typename T>
struct Vector2 {
    T x, y;
    Vector2() : x(0), y(0) {}
    Vector2(T X, T Y) : x(X), y(Y) {}
    template <typename U>
    explicit Vector2(const Vector2<U> &v) : x(static_cast<T>(v.x)), y(static_cast<T>(v.y)) {}
};
template <typename T> Vector2<T> operator+(const Vector2<T> &a, const Vector2<T> &b) { return Vector2<T>(a.x + b.x, a.y + b.y); }
template <typename T> Vector2<T> operator-(const Vector2<T> &a, const Vector2<T> &b) { return Vector2<T>(a.x - b.x, a.y - b.y); }
template <typename T> Vector2<T> operator*(const Vector2<T> &a, T s) { return Vector2<T>(a.x * s, a.y * s); }
template <typename T> Vector2<T> operator/(const Vector2<T> &a, T s) { return Vector2<T>(a.x / s, a.y / s); }
template <typename T> bool operator==(const Vector2<T> &a, const Vector2<T> &b) { return a.x == b.x && a.y == b.y; }
typedef Vector2<int> Vector2i;
typedef Vector2<unsigned int> Vector2u;
typedef Vector2<float> Vector2f;
}
