// Empty stand-in for Ygor's YgorMathIOOFF.h: WarpImages.cc includes it but uses nothing from it.
#pragma once
