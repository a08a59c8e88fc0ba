// Empty stand-in for the reference's Regex_Selectors.h: the Demons unit tests include it but use nothing from it.
#pragma once
