// stand-in: nothing from YgorString.h is used on the Demons path
#pragma once
