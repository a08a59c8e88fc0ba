// Empty stand-in for the reference's Thread_Pool.h: the Demons unit tests include it but use nothing from it.
#pragma once
