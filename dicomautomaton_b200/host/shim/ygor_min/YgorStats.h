// stand-in: nothing from YgorStats.h is used on the Demons path
#pragma once
