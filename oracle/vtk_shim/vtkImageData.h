/* stand-in for VTK's vtkImageData.h: see vtk_standin.h (test infrastructure) */
#include "vtk_standin.h"
