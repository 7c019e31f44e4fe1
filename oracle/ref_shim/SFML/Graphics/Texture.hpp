/* TEST INFRASTRUCTURE: stand-in for <SFML/Graphics/Texture.hpp>; include/util.hpp only names the type. */
#pragma once
#include <SFML/System/Vector2.hpp>
namespace sf {
class Texture {
public:
    Vector2u getSize() const { return Vector2u(0, 0); }
};
}
