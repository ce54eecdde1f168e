/* Stub of <alsa/control.h>: the SoapySX driver includes it (reference SoapySX.cpp:24) but
 * uses nothing from it. */
#pragma once
