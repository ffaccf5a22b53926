/* stand-in for VTK's vtkXMLImageDataReader.h: see vtk_standin.h (test infrastructure) */
#include "vtk_standin.h"
