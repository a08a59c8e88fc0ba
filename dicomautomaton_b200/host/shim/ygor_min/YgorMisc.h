// Minimal stand-in for Ygor's YgorMisc.h: isininc only.
#pragma once
template <class A, class B, class C>
inline bool isininc(A lo, B x, C hi) { return !(x < lo) && !(hi < x); }
