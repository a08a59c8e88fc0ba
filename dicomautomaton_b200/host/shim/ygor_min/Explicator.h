// Empty stand-in for Explicator.h (lexicon translation library): RegisterImagesDemons.cc includes it and uses nothing from it.
#pragma once
