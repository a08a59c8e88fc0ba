// Empty stand-in for the reference's Alignment_Rigid.h: the Demons unit tests include it but use nothing from it.
#pragma once
