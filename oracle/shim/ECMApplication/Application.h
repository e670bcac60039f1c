/* TEST INFRASTRUCTURE - empty stub: Environment.cpp:7 includes the GUI header without using it. */
#pragma once
