#pragma once
#include <type_traits>
namespace boost { template <class T> struct remove_pointer { typedef typename std::remove_pointer<T>::type type; }; }
