/* TEST INFRASTRUCTURE - stub of boost/polygon/voronoi.hpp (Boost 1.82 is not vendored by the
 * reference and not installed here).  Only `MedialAxis::VD` (ECMDataTypes.h:196-201) and
 * `ECM::Clear` (ECM.cpp:16) touch it on the translation units we build. */
#pragma once
namespace boost { namespace polygon {
template <class T> struct voronoi_diagram { void clear() {} };
}}
